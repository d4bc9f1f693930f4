"""Where does stream mode (K3 split / chunked K1b, K3a) start to pay?  Fixed total work (320 k frames, 16 kHz, level 13),
utterance length swept; per length the segment + spectrum stage times in utterance mode and in stream mode.
usage (GPU box): python profiles/sweep_stream_threshold.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from webspeechanalyzer_b200 import Engine, FaConfig, synth_speech  # noqa: E402

sr = 16000
base = [synth_speech(5 * sr, sr, 77, u) for u in range(64)]
rows = []
for secs in (5, 10, 20, 40, 80, 160, 320):
    n_utt = max(1, 8000 // secs)
    k = secs // 5
    pcms = [np.concatenate([base[(u * k + j) % 64] for j in range(k)]) for u in range(n_utt)]
    row = {"seconds": secs, "frames_per_utt": secs * 40, "utterances": n_utt}
    for name, env in (("utterance", {"FA_K3_MODE": "0", "FA_K1B_CHUNK": "0"}),
                      ("stream", {"FA_K3_MODE": "1"}),
                      ("stream_chunk1024", {"FA_K3_MODE": "1", "FA_K3_CHUNK": "1024", "FA_K3_WARM": "512", "FA_K1B_CHUNK": "1024"})):
        for kk in ("FA_K3_MODE", "FA_K3_CHUNK", "FA_K3_WARM", "FA_K1B_CHUNK"):
            os.environ.pop(kk, None)
        os.environ.update(env)
        with Engine(FaConfig.default(output_level=13)) as eng:
            eng.set_pipeline(1)
            for i, p in enumerate(pcms):
                eng.submit(i, p, sr)
            eng.upload(); eng.sync()
            for _ in range(3):
                eng.run_resident()
            eng.sync()
            acc = np.zeros(5)
            for _ in range(5):
                eng.run_resident(); eng.sync()
                st = eng.stage_times()
                acc += np.array([st[x] for x in ("spectrum", "peaks", "segment", "features", "total")])
            acc /= 5
            row[name] = {"spectrum_ms": acc[0], "segment_ms": acc[2], "features_ms": acc[3], "total_ms": acc[4],
                         "fixups": [eng.stream_fixups, eng.control_fixups]}
    rows.append(row)
    print(json.dumps(row), flush=True)
