"""Fixture for tests/test_predict.py::test_segment_vote_matches_reference_prediction_js: the reference's OWN
src/prediction.js (reset_predictions / predict_by_multiple_syllables / nn_prediction / seg_confidence_sort, lines 8-169)
executed by oracle/minijs on seeded classifier outputs, with `nn_mod.predict_single` replaced by a stub that hands the
callback the given ml5-shaped results (a flat class list for a one-row segment, a list of lists otherwise) and plotly / the DOM
stubbed out.  What the callback of predict_by_multiple_syllables received is stored: [top label, confidence / duration].
Run in the build container: python tests/golden/make_ref_js_vote_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
SRC = "/root/reference/src/prediction.js"
LABELS = ["N", "A", "S", "H"]


def reference_votes(segments):
    """segments: [(results, seg_time)] -> what the reference passed to callback_after_pred for each, in order."""
    from oracle.minijs.interp import Interp, JSArray, JSObject, Native, UNDEF
    src = open(SRC, encoding="utf-8").read()
    body = src[src.index("var sum_weights = 0;"):src.index("function plot_prediction_meters()")].replace("export function", "function")
    it = Interp()

    def to_js(v):
        if isinstance(v, dict):
            return JSObject({k: to_js(x) for k, x in v.items()})
        if isinstance(v, (list, tuple)):
            return JSArray([to_js(x) for x in v])
        if isinstance(v, (bool, str)) or v is None:
            return v
        return float(v)

    current = []

    def predict_single(this, args):      # (db, rows, callback, type, label): classify and call back -- here: the given results
        it.call(args[2], UNDEF, [to_js(current[0])])
        return UNDEF

    mod = it.call(it.eval_expression("(function(nn_mod, plot_prediction_meters, document){\n" + body +
                                     "\nreturn {reset_predictions: reset_predictions, predict_by_multiple_syllables: predict_by_multiple_syllables};})"),
                  UNDEF, [JSObject({"predict_single": Native(predict_single, "predict_single")}), Native(lambda t, a: UNDEF, "plot"), JSObject()])
    out = []

    def cb(this, args):
        out.append([args[1].a[0], args[1].a[1]])
        return UNDEF

    it.call(it.get(mod, "reset_predictions"), UNDEF, [True])
    for si, (res, seg_time) in enumerate(segments):
        current[:] = [res]
        n0 = len(out)
        it.call(it.get(mod, "predict_by_multiple_syllables"), UNDEF,
                ["cats", "emotion", float(si), to_js([[0.0] * 53] * len(seg_time)), to_js(seg_time), Native(cb, "cb")])
        it.run_microtasks()
        if len(out) == n0:
            out.append(None)             # seg_weight == 0: the reference never calls back
    return out


def make_cases(seed=3, n_clips=6):
    rng = np.random.default_rng(seed)
    clips = []
    for _ in range(n_clips):
        segs = []
        for _ in range(int(rng.integers(1, 7))):
            n_syl = int(rng.choice([1, 1, 2, 3, 5]))
            times = [[f"{rng.uniform(0, 9):.3f}", f"{rng.choice([0.0, 0.05, 0.2, 0.35, 1.1]) if rng.random() < 0.2 else rng.uniform(0.03, 0.9):.3f}"]
                     for _ in range(n_syl)]
            rows = []
            for _ in range(n_syl):
                p = rng.dirichlet(np.ones(4) * rng.choice([0.3, 1.0, 5.0]))
                if rng.random() < 0.15:
                    p = np.round(p, 1) / max(np.round(p, 1).sum(), 1e-9)      # exact ties between classes
                order = sorted(range(4), key=lambda i: -p[i])
                rows.append([{"label": LABELS[i], "confidence": float(p[i])} for i in order])
            segs.append({"results": rows[0] if n_syl == 1 else rows, "seg_time": times})
        clips.append(segs)
    return clips


def main():
    clips = make_cases()
    for segs in clips:
        votes = reference_votes([(s["results"], s["seg_time"]) for s in segs])
        for s, v in zip(segs, votes):
            s["reference_callback"] = v
    json.dump({"source": "src/prediction.js:8-169 executed by oracle/minijs", "clips": clips},
              open(os.path.join(HERE, "ref_js_vote.json"), "w"), indent=1)
    print(sum(len(c) for c in clips), "segments in", len(clips), "clips")


if __name__ == "__main__":
    main()
