import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from bench import make_workload, bench_config, SR
from webspeechanalyzer_b200 import Engine
cfg = bench_config(); pcms = make_workload(0, 1000)
eng = Engine(cfg)
spec_host = torch.empty((200000, 1024), dtype=torch.float32, pin_memory=True).numpy()
def step():
    t=[time.perf_counter()]
    eng.reset(); t.append(time.perf_counter())
    for i,p in enumerate(pcms): eng.submit(i,p,SR)
    t.append(time.perf_counter())
    eng.run(); t.append(time.perf_counter())
    eng.sync(); t.append(time.perf_counter())
    r = eng.result(None); t.append(time.perf_counter())
    n = eng._check(eng._lib.fa_copy_spectrum(eng._h, -1, spec_host.ctypes.data, 200000)); t.append(time.perf_counter())
    return np.diff(t)*1e3
for _ in range(3): step()
d = np.mean([step() for _ in range(5)], axis=0)
print('reset %.2f submit %.2f run(call) %.2f sync %.2f result %.2f spectrum %.2f ms total %.2f' % (*d, d.sum()))
