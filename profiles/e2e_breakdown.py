"""Host-side breakdown of the e2e step (two batches in flight): where the wall time of launch / wait / result goes.
usage (GPU box): python profiles/e2e_breakdown.py [depth] [sink|nosink] [f32|u8|feat]
  f32: the headline path (float32 PCM in, float32 dB rows out); u8: int16 PCM in, uint8 spectrum rows out; feat: int16 in, no spectrum"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_workload, bench_config, SR
from webspeechanalyzer_b200 import Engine

mode = sys.argv[3] if len(sys.argv) > 3 else "f32"
cfg = bench_config(); pcms = make_workload(0, 1000)
if mode == "u8":
    from webspeechanalyzer_b200 import FaConfig
    cfg = FaConfig.default(output_level=5, want_spectrum=1, spectrum_format=1)
elif mode == "feat":
    from webspeechanalyzer_b200 import FaConfig
    cfg = FaConfig.default(output_level=5, want_spectrum=0)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sink = (sys.argv[2] != "nosink") if len(sys.argv) > 2 else True
offs = np.zeros(1001, np.int64); offs[1:] = np.cumsum([p.size for p in pcms])
engs, specs, pcmh = [], [], []
cs = torch.cuda.Stream()
for j in range(depth):
    e = Engine(cfg); e.set_d2h_stream(cs.cuda_stream); engs.append(e)
    specs.append(torch.empty((200000, 1024), dtype=torch.uint8 if mode == "u8" else torch.float32, pin_memory=True).numpy())
    ph = torch.empty(int(offs[-1]), dtype=torch.float32 if mode == "f32" else torch.int16, pin_memory=True).numpy()
    for i, p in enumerate(pcms): ph[offs[i]:offs[i + 1]] = p if mode == "f32" else np.clip(np.rint(p * 32768.0), -32768, 32767).astype(np.int16)
    pcmh.append(ph)
T = {"reset": 0.0, "submit": 0.0, "run": 0.0, "sync": 0.0, "result": 0.0}
def launch(j):
    e = engs[j]
    t0 = time.perf_counter(); e.reset(); t1 = time.perf_counter()
    e.submit_batch(0, pcmh[j], offs, SR); (e.set_spectrum_sink(specs[j] if sink else None) if mode != "feat" else None); t2 = time.perf_counter()
    e.run(); t3 = time.perf_counter()
    T["reset"] += t1 - t0; T["submit"] += t2 - t1; T["run"] += t3 - t2
def collect(j):
    e = engs[j]
    t0 = time.perf_counter(); e.sync(); t1 = time.perf_counter(); r = e.result(None); t2 = time.perf_counter()
    T["sync"] += t1 - t0; T["result"] += t2 - t1
def steps(k):
    fl = []
    for i in range(k):
        j = i % depth
        if len(fl) == depth: collect(fl.pop(0))
        launch(j); fl.append(j)
    while fl: collect(fl.pop(0))
steps(2 * depth)
for k in T: T[k] = 0.0
K = 10
t0 = time.perf_counter(); steps(K); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"mode {mode} depth {depth} sink {sink}: {1e3*dt/K:.2f} ms/step; host ms/step: " + ", ".join(f"{k} {1e3*v/K:.2f}" for k, v in T.items()))
