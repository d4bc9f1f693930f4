"""Per-stage CUDA-event times of the C2 batch (1000 x 5 s x 16 kHz, level 5 + spectrum), serial mode, for a list of
environment variants.  usage: python profiles/stage_times.py [name=ENV1=v,ENV2=v ...]  -> one JSON line per variant."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from concurrent.futures import ThreadPoolExecutor  # noqa: E402

from webspeechanalyzer_b200 import Engine, FaConfig, synth_speech  # noqa: E402

SR, N = 16000, int(os.environ.get("N_UTT", "1000"))
LEVEL = int(os.environ.get("LEVEL", "5"))
with ThreadPoolExecutor(16) as ex:
    pcms = list(ex.map(lambda u: synth_speech(5 * SR, SR, 20261017, u), range(N)))
variants = sys.argv[1:] or ["default="]
for v in variants:
    name, _, envs = v.partition("=")
    sets = [kv.split(":") for kv in envs.split(",") if kv]
    old = {k: os.environ.get(k) for k, _ in sets}
    for k, val in sets:
        os.environ[k] = val
    cfg = FaConfig.default(output_level=LEVEL, want_spectrum=int(os.environ.get("WANT_SPEC", "1")))
    e = Engine(cfg)
    e.set_pipeline(1)
    for i, p in enumerate(pcms):
        e.submit(i, p, SR)
    e.upload()
    acc = np.zeros(5)
    for _ in range(3):
        e.run_resident()
    reps = 10
    for _ in range(reps):
        e.run_resident()
        e.sync()
        st = e.stage_times()
        acc += np.array([st[k] for k in ("spectrum", "peaks", "segment", "features", "total")])
    e.download(); e.sync()
    c = e.counts()
    out = {"variant": name, "env": dict(sets), **{k: round(float(x) / reps, 4) for k, x in zip(("spectrum", "peaks", "segment", "features", "total"), acc)},
           "segments": c["segments"], "feature_rows": c["feature_rows"], "overflow": c["overflow"], "k3_redos": e.k3_redos,
           "launches": e.launches}
    print(json.dumps(out), flush=True)
    e.close()
    for k, val in old.items():
        if val is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = val
