"""What bounds the uint8-spectrum e2e step (int16 PCM in, uint8 rows out)?  Sweep of batches in flight x sub-batches per batch,
same call sequence as bench.py's e2e_byte_spectrum.  usage (GPU box): python profiles/e2e_u8_sweep.py"""
import os, sys, time, json
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_workload, SR
from webspeechanalyzer_b200 import Engine, FaConfig

pcms = make_workload(0, 1000)
offs = np.zeros(1001, np.int64); offs[1:] = np.cumsum([p.size for p in pcms])
i16 = np.concatenate([np.clip(np.rint(p * 32768.0), -32768, 32767).astype(np.int16) for p in pcms])
for fetch in ("all", "features"):
    for depth in (2, 3):
        for subs in (0, 2, 4, 8):
            cfg = FaConfig.default(output_level=5, want_spectrum=1, spectrum_format=1)
            engs, sinks, hosts, streams = [], [], [], []
            for j in range(depth):
                st = torch.cuda.Stream(); streams.append(st)
                e = Engine(cfg); e.set_stream(st.cuda_stream); e.set_pipeline(subs); engs.append(e)
                sinks.append(torch.empty((200000, 1024), dtype=torch.uint8, pin_memory=True).numpy())
                ph = torch.empty(int(offs[-1]), dtype=torch.int16, pin_memory=True).numpy(); ph[:] = i16; hosts.append(ph)
            def launch(j):
                engs[j].reset(); engs[j].submit_batch(0, hosts[j], offs, SR); engs[j].set_spectrum_sink(sinks[j]); engs[j].run()
            def collect(j):
                engs[j].sync()
                if fetch == "all": engs[j].result(None)
                else: engs[j].feature_table()
            def steps(k):
                fl = []
                for i in range(k):
                    j = i % depth
                    if len(fl) == depth: collect(fl.pop(0))
                    launch(j); fl.append(j)
                while fl: collect(fl.pop(0))
            steps(2 * depth)
            K = 12
            t0 = time.perf_counter(); steps(K); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(json.dumps({"fetch": fetch, "depth": depth, "sub_batches": subs, "ms_per_step": round(1e3 * dt / K, 3)}), flush=True)
            for e in engs: e.close()
            del sinks, hosts
