import sys, numpy as np
sys.path.insert(0, '/root/repo')
from webspeechanalyzer_b200 import Engine, FaConfig, synth_speech
sr = 48000
minute = np.concatenate([synth_speech(5 * sr, sr, 4, u) for u in range(12)])
stream = np.tile(minute, 10)
cfg = FaConfig.default(output_level=13)
with Engine(cfg) as eng:
    eng.submit(0, stream, sr)
    eng.run(); eng.sync()
    print(eng.counts())
