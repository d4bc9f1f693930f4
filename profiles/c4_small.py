"""A stream as ONE utterance (for ncu launch lists / fix-up counts): python profiles/c4_small.py [minutes]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from webspeechanalyzer_b200 import Engine, FaConfig, synth_speech  # noqa: E402

minutes = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sr = 48000
minute = np.concatenate([synth_speech(5 * sr, sr, 4, u) for u in range(12)])
stream = np.tile(minute, minutes)
cfg = FaConfig.default(output_level=13)
with Engine(cfg) as eng:
    eng.set_pipeline(1)
    eng.submit(0, stream, sr)
    eng.run(); eng.sync()
    eng.upload(); eng.sync()
    eng.run_resident(); eng.sync()
    print(eng.counts(), "stage ms", eng.stage_times(), "fixups: smoothing", eng.stream_fixups, "control", eng.control_fixups)
