import sys, time; sys.path.insert(0, '.')
import numpy as np
from webspeechanalyzer_b200 import FaConfig
from webspeechanalyzer_b200.engine import Engine, synth_speech
from oracle import oracle

def check(cfg, pcms, sr, name):
    eng = Engine(cfg)
    for i, p in enumerate(pcms): eng.submit(i, p, sr)
    eng.run(); eng.sync()
    print(name, 'stage ms', eng.stage_times(), 'launches', eng.launches)
    ok = True
    for i, p in enumerate(pcms):
        fe = oracle.frontend(cfg, p, sr, spectrum=True, frames=True)
        fr = eng.frames(i)
        eq = np.array_equal(fr, fe['frames'])
        nd = int((fr != fe['frames']).sum())
        msg = f'  utt{i}: frames bit-exact={eq} (ndiff {nd}/{fr.size})'
        if cfg.want_spectrum or cfg.output_level <= 2:
            sp = eng.spectrum(i)
            fin = np.isfinite(fe['spectrum']) & np.isfinite(sp)
            msg += f' spectrum max|d|={np.abs(sp[fin]-fe["spectrum"][fin]).max():.2e} dB'
        if cfg.output_level >= 3:
            an = oracle.analyze_frames(cfg, fe['frames'])
            r = eng.result(i)
            msg += f' segs {r.seg_ci == an.seg_ci} ({len(an.seg_ci)})'
            msg += f' formants {np.array_equal(r.formants, an.formants)} energy {np.array_equal(r.energy, an.energy)}'
            msg += f' syl {np.array_equal(r.syllables, an.syllables)} ({len(an.syllables)})'
            msg += f' feat {r.features.shape == an.features.shape and np.array_equal(r.features, an.features, equal_nan=True)} ({an.features.shape[0]})'
            if r.features.shape == an.features.shape and an.features.size:
                with np.errstate(all='ignore'):
                    rel = np.abs(r.features-an.features)/np.maximum(np.abs(an.features),1e-300)
                msg += f' maxrel {np.nanmax(rel):.2e}'
            if r.seg_ci != an.seg_ci: print(r.seg_ci, an.seg_ci)
            ok &= r.seg_ci == an.seg_ci
        ok &= eq
        print(msg)
    eng.close()
    return ok

sr = 16000
pcms = [synth_speech(5*sr, sr, 1, u) for u in range(6)]
ok = True
for lvl in (13, 5, 4, 10):
    ok &= check(FaConfig.default(output_level=lvl, want_spectrum=1 if lvl==13 else 0), pcms, sr, f'level{lvl} 16k')
ok &= check(FaConfig.default(output_level=13, window_step_ms=15.0), [synth_speech(3*44100, 44100, 2, u) for u in range(3)], 44100, 'level13 44.1k step15')
ok &= check(FaConfig.default(output_level=13, fft_size=1024, smoothing=0.0), pcms[:2], sr, 'N1024 tau0')
ok &= check(FaConfig.default(output_level=5, fft_size=4096, spec_type=3), pcms[:2], sr, 'N4096 dfft')
ok &= check(FaConfig.default(output_level=2), pcms[:2], sr, 'spectrum only')
print('ALL OK' if ok else 'FAILURES')
