"""BASELINE config 5: fftSize sweep 512..8192 x smoothingTimeConstant {0, 0.8} on the C2 shape (1000 x 5 s x 16 kHz),
per-stage device time and fraction of the measured HBM peak from the stage's algorithmic bytes.
usage (GPU box): python profiles/sweep_c5.py [n_utt] > gpurun_out/sweep_c5.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, measured_peaks, SR, SECONDS  # noqa: E402
from webspeechanalyzer_b200 import Engine, FaConfig  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pcms = make_workload(0, n_utt)
peak, src = measured_peaks()
rows = []
for N in (512, 1024, 2048, 4096, 8192):
    for tau in (0.0, 0.8):
        cfg = FaConfig.default(output_level=5, want_spectrum=1, fft_size=N, smoothing=tau)
        eng = Engine(cfg)
        eng.set_pipeline(1)
        for i, p in enumerate(pcms):
            eng.submit(i, p, SR)
        eng.upload()
        for _ in range(3):
            eng.run_resident()
        acc = np.zeros(5)
        K = 5
        for _ in range(K):
            eng.run_resident()
            eng.sync()
            st = eng.stage_times()
            acc += np.array([st[k] for k in ("spectrum", "peaks", "segment", "features", "total")])
        acc /= K
        eng.download(); eng.sync()
        tot = eng.counts()
        F, B, hop = tot["frames"], 128, 400
        alg = {"spectrum": F * (4 * hop + 4 * (N // 2) + 4 * B), "peaks": F * (4 * B + 36), "segment": F * (4 * B + 36) + tot["formant_rows"] * 48,
               "features": tot["formant_rows"] * 36 + tot["feature_rows"] * 424}
        row = {"fft_size": N, "tau": tau, "audio_s_per_s": n_utt * SECONDS / (acc[4] * 1e-3), "ms": {}, "frac_hbm": {},
               "kernel_path": "fast (K1a+K1b)" if N == 2048 else "generic shared-memory radix-2" if N < 2048 else
               "big: register path for 10 stages + shared-memory tail"}
        for i, k in enumerate(("spectrum", "peaks", "segment", "features")):
            row["ms"][k] = float(acc[i])
            row["frac_hbm"][k] = alg[k] / (acc[i] * 1e-3) / 1e9 / peak if acc[i] > 0 else None
        rows.append(row)
        eng.close()
print(json.dumps({"peak_gbs": peak, "peak_source": src, "n_utt": n_utt, "rows": rows}, indent=1))
