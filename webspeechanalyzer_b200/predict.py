"""Emotion / label prediction on the 53-dim rows, the way the web app does it.

    reference                                                          here
    ml5.neuralNetwork(...).load(model.json, model_meta.json, weights)  load_tfjs_model(dir)  (tf.js layers format)
    nn.classifyMultiple(rows, cb)   src/neuralmodel.js:540-585          Classifier.classify_multiple(rows)
    predict_by_multiple_syllables / nn_prediction / seg_confidence_sort  SegmentVoter
                                    src/prediction.js:47-169

The forward pass runs on the GPU (csrc/fa_mlp.cu behind fa_mlp_* of include/fa_b200.h); there is no CPU path.
Classifier.classify_features(engine) classifies the rows of a finished batch where they are -- in HBM.
Model files are read at run time from a directory the caller names (e.g. dist/nnmodel/1/cats_emotion of the web app);
nothing of the reference's models is stored in this repository."""
from __future__ import annotations

import ctypes as C
import json
import math
import os

import numpy as np

from . import _capi
from ._capi import FaError

ACTIVATIONS = {"linear": 0, "relu": 1, "sigmoid": 2, "softmax": 3}


def load_tfjs_model(model_dir: str, model_json: str = "model.json", meta_json: str = "model_meta.json") -> dict:
    """tf.js layers-model (Sequential of Dense) + ml5 meta -> dict(dims, activations, kernels, biases, in_min, in_max, labels)."""
    with open(os.path.join(model_dir, model_json)) as f:
        doc = json.load(f)
    with open(os.path.join(model_dir, meta_json)) as f:
        meta = json.load(f)
    topo = doc["modelTopology"]
    layers = (topo.get("config") or topo["model_config"]["config"])["layers"]
    dense = [l for l in layers if l["class_name"] == "Dense"]
    if len(dense) != len(layers):
        raise ValueError("only Sequential models of Dense layers are supported")
    # weights: manifest order, float32 little-endian, concatenated over the listed files
    blobs, specs = b"", []
    for group in doc["weightsManifest"]:
        for p in group["paths"]:
            with open(os.path.join(model_dir, p), "rb") as f:
                blobs += f.read()
        specs += group["weights"]
    arrays, off = {}, 0
    for w in specs:
        if w["dtype"] != "float32":
            raise ValueError("only float32 weights are supported")
        n = int(np.prod(w["shape"])) if w["shape"] else 1
        arrays[w["name"]] = np.frombuffer(blobs, "<f4", n, off).reshape(w["shape"]).copy()
        off += 4 * n
    kernels, biases, acts, dims = [], [], [], []
    for l in dense:
        name = l["config"]["name"]
        k = arrays[name + "/kernel"]
        b = arrays.get(name + "/bias")
        if b is None:
            b = np.zeros(k.shape[1], np.float32)
        kernels.append(np.ascontiguousarray(k, np.float32))
        biases.append(np.ascontiguousarray(b, np.float32))
        acts.append(ACTIVATIONS[l["config"]["activation"]])
        if not dims:
            dims.append(k.shape[0])
        dims.append(k.shape[1])
    nin = dims[0]
    ins = meta["inputs"]
    if meta.get("isNormalized", True):
        in_min = np.array([ins[str(i)]["min"] for i in range(nin)], np.float64)
        in_max = np.array([ins[str(i)]["max"] for i in range(nin)], np.float64)
    else:   # ml5 normalises a row only when the saved meta says the training data was (NeuralNetwork.classifyInternal)
        in_min, in_max = np.zeros(nin, np.float64), np.ones(nin, np.float64)
    labels = None
    outs = meta.get("outputs") or {}
    for o in outs.values():
        if "legend" in o:       # one-hot legend: label -> vector
            labels = [None] * len(o["legend"])
            for lab, vec in o["legend"].items():
                labels[int(np.argmax(vec))] = lab
    return dict(dims=dims, activations=acts, kernels=kernels, biases=biases, in_min=in_min, in_max=in_max,
                labels=labels or [str(i) for i in range(dims[-1])])


class Classifier:
    """A loaded model on one GPU."""

    def __init__(self, model: dict | str, device: int = 0):
        if isinstance(model, str):
            model = load_tfjs_model(model)
        self.model = model
        self.labels = list(model["labels"])
        self._lib = _capi.lib()
        n = len(model["kernels"])
        dims = (C.c_int * (n + 1))(*model["dims"])
        acts = (C.c_int * n)(*model["activations"])
        self._keep = [np.ascontiguousarray(k, np.float32) for k in model["kernels"]] + \
                     [np.ascontiguousarray(b, np.float32) for b in model["biases"]]
        kp = (C.c_void_p * n)(*[a.ctypes.data for a in self._keep[:n]])
        bp = (C.c_void_p * n)(*[a.ctypes.data for a in self._keep[n:]])
        lo = np.ascontiguousarray(model["in_min"], np.float64)
        hi = np.ascontiguousarray(model["in_max"], np.float64)
        self._h = C.c_void_p()
        rc = self._lib.fa_mlp_create(n, dims, acts, kp, bp, lo.ctypes.data, hi.ctypes.data, device, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise FaError(rc, self._lib.fa_status_string(rc).decode())
        self.n_in, self.n_out = model["dims"][0], model["dims"][-1]

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.fa_mlp_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise FaError(rc, (self._lib.fa_mlp_last_error(self._h) or b"").decode() or self._lib.fa_status_string(rc).decode())
        return rc

    def probabilities(self, rows) -> np.ndarray:
        rows = np.ascontiguousarray(rows, np.float64).reshape(-1, self.n_in)
        out = np.empty((rows.shape[0], self.n_out), np.float32)
        self._check(self._lib.fa_mlp_classify(self._h, rows.ctypes.data, rows.shape[0], out.ctypes.data))
        return out

    def classify_features(self, engine, utt_id: int | None = None) -> np.ndarray:
        """Class scores of the feature rows of a finished Engine batch, computed where the rows are (device)."""
        c = engine.counts(utt_id)
        out = np.empty((c["feature_rows"], self.n_out), np.float32)
        n = self._check(self._lib.fa_mlp_classify_features(self._h, engine._h, -1 if utt_id is None else utt_id,
                                                           out.ctypes.data, out.shape[0]))
        return out[:n]

    def results(self, probs: np.ndarray) -> list:
        """ml5 classifyMultiple's shape: per row a list of {label, confidence}, highest confidence first."""
        out = []
        for p in probs:
            order = sorted(range(len(p)), key=lambda i: -float(p[i]))      # stable: ties keep label order
            out.append([{"label": self.labels[i], "confidence": float(p[i])} for i in order])
        return out

    def classify_multiple(self, rows) -> list:
        return self.results(self.probabilities(rows))


class SegmentVoter:
    """predict_by_multiple_syllables + nn_prediction + seg_confidence_sort (src/prediction.js:47-169) for one or more
    models ("DBs"): sqrt(duration)-weighted confidence sums per label, per segment and over the whole clip."""

    def __init__(self, db_ids=(1,)):
        self.db_ids = list(db_ids)
        self.conf_all = {d: {} for d in self.db_ids}
        self.sum_weights = 0.0
        self.max_inv_entropy, self.min_entropy_db = 0.0, None

    def segment(self, results_by_db: dict, seg_time) -> tuple | None:
        """results_by_db[db] = ml5-shaped results of the segment's syllable rows; seg_time = [[t0, dur] strings ...].
        Returns (top label, confidence / total duration) like the callback of predict_by_multiple_syllables."""
        seg_weight = 0.0
        for t in seg_time:          # a plain left-to-right sum like the reference's loop (Python >= 3.12's sum() compensates)
            seg_weight += float(t[1])
        if not seg_weight > 0:
            return None
        conf_seg = {d: {} for d in self.db_ids}
        for d in self.db_ids:
            res = results_by_db[d]
            for ph, t in enumerate(seg_time):
                w = math.sqrt(float(t[1]))
                # with a single syllable the app reads result_out[ph] of the FLAT class list: only the top class counts
                items = [res[0][0]] if len(seg_time) == 1 else res[ph]
                for it in items:
                    wc = it["confidence"] * w
                    self.conf_all[d][it["label"]] = self.conf_all[d].get(it["label"], 0) + wc if self.conf_all[d].get(it["label"]) else wc
                    conf_seg[d][it["label"]] = conf_seg[d].get(it["label"], 0) + wc if conf_seg[d].get(it["label"]) else wc
        self.sum_weights += seg_weight
        best, top = 0.0, None
        for d in self.db_ids:
            max_all, max_seg, lab_seg = 0.0, 0.0, None
            for lab in self.conf_all[d]:
                if self.conf_all[d][lab] > max_all:
                    max_all = self.conf_all[d][lab]
                if conf_seg[d].get(lab, 0) > max_seg:
                    max_seg, lab_seg = conf_seg[d][lab], lab
            if max_seg > best:
                best, top = max_seg, lab_seg
            if max_all > self.max_inv_entropy:
                self.max_inv_entropy, self.min_entropy_db = max_all, d
        return top, best / seg_weight
