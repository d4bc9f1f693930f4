#!/usr/bin/env python
"""BASELINE.json configs 3 and 4 at (per-GPU) size on one B200, through the C-ABI, checked against the CPU oracle.

C3: 48 kHz 5 s utterances, Syllable Features (output_level 13), feature-only (no spectrum output) -- `--c3-utts` per GPU
    (the 100k-utterance config is 12 500 per GPU on 8 GPUs = 12 GB of float32 PCM; utterances are generated on the host
    and repeated to reach the count, so the host does not spend minutes synthesising).
C4: one continuous 1-hour 48 kHz stream (172.8 M samples, 144 000 frames) as ONE utterance: K1a frame-parallel, K1b one
    CTA and K3 one warp carry the exact smoothing / segmentation state across the whole hour (no stitching).
Prints one JSON object; wall-clock around fa_run .. fa_sync with host PCM (pinned staging inside the library)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from webspeechanalyzer_b200 import Engine, FaConfig, synth_speech  # noqa: E402


def timed(eng, reps=3):
    ts = []
    for _ in range(reps):
        eng.upload()
        eng.sync()                    # the H2D copy is asynchronous: keep it out of the resident timing
        t0 = time.perf_counter()
        eng.run_resident()
        eng.sync()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-utts", type=int, default=12500)
    ap.add_argument("--c4-seconds", type=int, default=3600)
    ap.add_argument("--check", type=int, default=1)
    a = ap.parse_args()
    out = {}
    sr = 48000
    # ---------------- C3
    cfg = FaConfig.default(output_level=13)
    base = [synth_speech(5 * sr, sr, 20261017, u) for u in range(250)]
    with Engine(cfg) as eng:
        t0 = time.perf_counter()
        for u in range(a.c3_utts):
            eng.submit(u, base[u % len(base)], sr)
        t_submit = time.perf_counter() - t0
        t0 = time.perf_counter()
        eng.run(); eng.sync()
        t_e2e = time.perf_counter() - t0
        c = eng.counts()
        t_res = timed(eng)
        st = eng.stage_times()
        audio = a.c3_utts * 5.0
        out["C3"] = {"utterances": a.c3_utts, "sample_rate": sr, "level": 13, "pcm_GB": a.c3_utts * 5 * sr * 4 / 1e9,
                     "frames": c["frames"], "syllable_rows": c["feature_rows"], "segments": c["segments"],
                     "resident_s": t_res, "resident_audio_s_per_s": audio / t_res, "resident_frames_per_s": c["frames"] / t_res,
                     "host_run_to_sync_s": t_e2e, "host_audio_s_per_s": audio / t_e2e, "submit_memcpy_s": t_submit,
                     "stage_ms_pipelined_run": st, "overflow": c["overflow"]}
        if a.check:
            from oracle import oracle
            ok = True
            for u in (0, 249, a.c3_utts - 1):
                fe, an = oracle.analyze_pcm(cfg, base[u % len(base)], sr)
                r = eng.result(u)
                ok = ok and r.seg_ci == an.seg_ci and np.array_equal(r.features, an.features, equal_nan=True) \
                    and np.array_equal(r.syllables, an.syllables)
            out["C3"]["parity_vs_oracle_sampled"] = bool(ok)
    # ---------------- C4
    minute = np.concatenate([synth_speech(5 * sr, sr, 4, u) for u in range(12)])
    stream = np.tile(minute, a.c4_seconds // 60)
    cfg = FaConfig.default(output_level=13)
    with Engine(cfg) as eng:
        eng.submit(0, stream, sr)
        t0 = time.perf_counter()
        eng.run(); eng.sync()
        t_e2e = time.perf_counter() - t0
        c = eng.counts()
        t_res = timed(eng, reps=2)
        eng.set_pipeline(1)
        eng.upload(); eng.run_resident(); eng.sync()
        st = eng.stage_times()
        out["C4"] = {"seconds": a.c4_seconds, "samples": int(stream.size), "frames": c["frames"], "segments": c["segments"],
                     "syllable_rows": c["feature_rows"], "resident_s": t_res, "x_realtime_resident": a.c4_seconds / t_res,
                     "host_run_to_sync_s": t_e2e, "x_realtime_host": a.c4_seconds / t_e2e, "stage_ms_serial": st,
                     "overflow": c["overflow"]}
        if a.check:
            from oracle import oracle
            t0 = time.perf_counter()
            fe, an = oracle.analyze_pcm(cfg, stream, sr)
            out["C4"]["oracle_cpu_s"] = time.perf_counter() - t0
            r = eng.result(0)
            out["C4"]["parity_vs_oracle"] = bool(r.seg_ci == an.seg_ci and np.array_equal(r.features, an.features, equal_nan=True)
                                                 and np.array_equal(r.syllables, an.syllables) and np.array_equal(r.formants, an.formants))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
