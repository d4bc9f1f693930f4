"""Regenerates tests/golden/ref_js_api.json: what the reference's PUBLIC API module (inner module 1 of the formantanalyzer
bundle, /root/reference/dist/main.js:2@B2714 -- configure / LaunchAudioNodes) does, executed by oracle/minijs with its
browser-only audio-graph module replaced by a recording stub (oracle/minijs/run_reference.py: ReferenceAPI).

    python tests/golden/make_ref_js_api_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.minijs.run_reference import ReferenceAPI  # noqa: E402

FULL = dict(plot_enable=False, spec_type=1, output_level=4, plot_len=200, f_min=50, f_max=4000, N_fft_bins=256, N_mel_bins=128,
            window_width=25, window_step=25, pause_length=200, min_seg_length=50, auto_noise_gate=True, voiced_max_dB=100,
            voiced_min_dB=10, plot_lag=1, pre_norm_gain=1000, high_f_emph=0)

CONFIGURE_CASES = [
    dict(FULL),
    dict(FULL, spec_type=3, output_level=0, f_min=0, f_max=8000, N_fft_bins=0, N_mel_bins=64, window_width=0, window_step=15,
         pause_length=0, min_seg_length=100, auto_noise_gate=False, voiced_max_dB=0, voiced_min_dB=0, pre_norm_gain=0),
    dict(FULL, output_level=13, window_step=15, high_f_emph=0.05, voiced_min_dB=30, spec_type=2, N_fft_bins=128),
    dict(FULL, output_level=5, plot_enable=True, plot_canvas=None, plot_len=80),          # plot_enable without a canvas: stays off
    dict(FULL, output_level=11, f_max=0, pause_length=120, min_seg_length=0),
    {"output_level": 13, "window_step": 15},                                                # partial object: see the test
]

LAUNCH_CASES = [   # (name, playing, args)
    ("file_online", False, [1, "buffer", None, ["lab"], False, False]),
    ("file_offline", False, [1, "buffer", None, [], True, True, 0.5, 2.0]),
    ("file_without_source", False, [1, None]),
    ("element", False, [2, "audio-element"]),
    ("microphone", False, [3]),
    ("unknown_source", False, [4, {"pcm": [0.0], "sampleRate": 16000}]),
    ("already_playing", True, [1, "buffer"]),
]


def main():
    doc = {"generator": "tests/golden/make_ref_js_api_golden.py", "configure": [], "launch": []}
    for cfg in CONFIGURE_CASES:
        A = ReferenceAPI()
        doc["configure"].append({"cfg": cfg, "settings": A.configure(cfg)})
    for name, playing, args in LAUNCH_CASES:
        A = ReferenceAPI()
        A.configure(dict(FULL, output_level=13, window_step=15))
        A.playing = playing
        status, value = A.launch(*args)
        doc["launch"].append({"name": name, "playing": playing, "args": args, "status": status, "value": value,
                              "calls": [[c[0], c[1]] for c in A.calls]})
    A = ReferenceAPI()
    A.fail["reset_nodes"] = "Invalid reset_nodes config"
    status, value = A.launch(1, "buffer")
    doc["launch"].append({"name": "reset_nodes_rejects", "playing": False, "args": [1, "buffer"], "status": status, "value": value,
                          "calls": [[c[0], c[1]] for c in A.calls]})
    with open(os.path.join(HERE, "ref_js_api.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote ref_js_api.json:", len(doc["configure"]), "configure cases,", len(doc["launch"]), "launch cases")


if __name__ == "__main__":
    main()
