"""One fft size of the C5 sweep, serial pipeline (for ncu launch lists): python profiles/c5_one.py <fft_size> [n_utt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_workload, SR  # noqa: E402
from webspeechanalyzer_b200 import Engine, FaConfig  # noqa: E402

N = int(sys.argv[1])
n_utt = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
pcms = make_workload(0, n_utt)
cfg = FaConfig.default(output_level=5, want_spectrum=1, fft_size=N, smoothing=0.8)
with Engine(cfg) as eng:
    eng.set_pipeline(1)
    for i, p in enumerate(pcms):
        eng.submit(i, p, SR)
    eng.upload()
    for _ in range(3):
        eng.run_resident()
    eng.sync()
    print(N, eng.stage_times())
