"""Kernel timeline of the resident region (two batches in flight): CUDA-event time stamps of every sub-batch's stage
boundaries on a common origin.  usage (GPU box): FA_TRACE=1 python profiles/resident_timeline.py [depth] [sub_batches]"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ["FA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_workload, bench_config, SR
from webspeechanalyzer_b200 import Engine

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
subs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = bench_config(); pcms = make_workload(0, 1000)
engs = []
for j in range(depth):
    e = Engine(cfg); st = torch.cuda.Stream(); e.set_stream(st.cuda_stream); e._st = st; e.set_pipeline(subs)
    for i, p in enumerate(pcms): e.submit(i, p, SR)
    e.upload(); engs.append(e)
for k in range(6 * depth): engs[k % depth].run_resident()
torch.cuda.synchronize()
for k in range(4 * depth): engs[k % depth].run_resident()
torch.cuda.synchronize()
for e in engs: e.sync()      # prints the stamps of each handle's LAST run
