"""Host-side mirror of the formantanalyzer public API (the drop-in boundary).

    reference (inner module 1 of /root/reference/dist/main.js line 2)           here
    configure(cfg)                                   @B3292                      configure(cfg)
    LaunchAudioNodes(context_source, source_obj, callback, file_labels=[],
                     offline=false, test_play=true, play_offset, play_duration)  LaunchAudioNodes(...)  -> Future(True)
                                                     @B4469
    StopAudioNodes(reason)                           @B5699                      StopAudioNodes(reason)
    set_predicted_label_for_segment(si, idx, label)  (module 2 `T` -> set_segments_label @B31782-)  same name

The reference's host language is JavaScript; `node` is not present in this image, so the tested host shim
is this Python module over the C-ABI (ctypes), and the N-API addon that binds the same C-ABI for Node lives
in webspeechanalyzer_b200/node/ (compile-checked only).  Same names, argument order, callback
shapes per output level (P() @B28869), string rejections and single-flight rule; callbacks are suppressed
when test_play is true (the reference's default, quirk 13).

Extensions (documented in INTEGRATION.md): context_source 4 = {"pcm": float32 array, "sampleRate": sr};
configure() accepts fftSize, smoothingTimeConstant, minDecibels, maxDecibels, mag_scale, clamp_dB;
levels 1-2 (canvas only in the reference) call back once with the dB spectrum.
"""
from __future__ import annotations

import math
from concurrent.futures import Future
from decimal import Decimal, ROUND_HALF_UP

import numpy as np

from . import wav
from ._ctypes_defs import FaConfig
from .engine import Engine

__all__ = ["configure", "LaunchAudioNodes", "StopAudioNodes", "set_predicted_label_for_segment", "LaunchError", "LiveSession"]


class LaunchError(Exception):
    """Carries the string the reference's promise would have been rejected with."""


def _defaults() -> dict:
    # /root/reference/dist/main.js:2@B2972
    return dict(plot_enable=False, spec_type=1, output_level=4, plot_len=200, f_min=50, f_max=4000, N_fft_bins=256,
                N_mel_bins=128, window_width=25, window_step=25, pause_length=200, min_seg_length=50,
                auto_noise_gate=True, voiced_max_dB=100, voiced_min_dB=10, plot_lag=1, pre_norm_gain=1000,
                high_f_emph=0, plot_canvas=None, canvas_width=200, canvas_height=100,
                # extension fields (AnalyserNode front end)
                fftSize=2048, smoothingTimeConstant=0.8, minDecibels=-100, maxDecibels=-30, mag_scale=0, clamp_dB=True,
                device=0)


_settings = _defaults()
_state = {"playing": False, "stop": None, "engine": None, "engine_key": None, "labels": []}


def reset_defaults():
    """Test helper: module state back to the formantanalyzer defaults."""
    global _settings
    _settings = _defaults()
    _state.update(playing=False, stop=None, labels=[])


def configure(cfg: dict) -> None:
    """configure() @B3292: `x && (a.x = x)` for the truthy-tested fields, `null !== x` for the others."""
    a = _settings
    truthy = ("output_level", "f_max", "N_fft_bins", "N_mel_bins", "window_width", "window_step", "pre_norm_gain",
              "pause_length", "min_seg_length", "voiced_max_dB")
    not_null = ("spec_type", "f_min", "high_f_emph", "auto_noise_gate", "voiced_min_dB")
    for k in truthy:
        if cfg.get(k):
            a[k] = cfg[k]
    for k in not_null:
        if k in cfg and cfg[k] is not None:
            a[k] = cfg[k]
    if cfg.get("plot_enable") and cfg.get("plot_canvas"):
        # plotting is out of scope (canvas UI); remember the fields the segmentor depends on
        a["plot_enable"] = True
        if cfg.get("plot_len"):
            a["plot_len"] = cfg["plot_len"]
    else:
        a["plot_enable"] = False
    for k in ("fftSize", "smoothingTimeConstant", "minDecibels", "maxDecibels", "mag_scale", "clamp_dB", "device"):
        if k in cfg and cfg[k] is not None:
            a[k] = cfg[k]


def _fa_config() -> FaConfig:
    a = _settings
    return FaConfig.default(
        spec_type=int(a["spec_type"]), output_level=int(a["output_level"]), plot_len=int(a["plot_len"]),
        n_fft_bins=int(a["N_fft_bins"]), n_mel_bins=int(a["N_mel_bins"]), auto_noise_gate=1 if a["auto_noise_gate"] else 0,
        f_min=float(a["f_min"]), f_max=float(a["f_max"]), window_width_ms=float(a["window_width"]),
        window_step_ms=float(a["window_step"]), pause_length_ms=float(a["pause_length"]),
        min_seg_length_ms=float(a["min_seg_length"]), voiced_max_db=float(a["voiced_max_dB"]),
        voiced_min_db=float(a["voiced_min_dB"]), pre_norm_gain=float(a["pre_norm_gain"]),
        high_f_emph=float(a["high_f_emph"]), fft_size=int(a["fftSize"]), clamp_db=1 if a["clamp_dB"] else 0,
        smoothing=float(a["smoothingTimeConstant"]), min_db=float(a["minDecibels"]), max_db=float(a["maxDecibels"]),
        mag_scale=float(a["mag_scale"]))


def to_fixed3(x: float) -> str:
    """Number.prototype.toFixed(3) (get_syls_timestamps @B31114): exact decimal value, ties up."""
    return str(Decimal(float(x)).quantize(Decimal("0.001"), rounding=ROUND_HALF_UP))


def track_arrays(tracks, points) -> list[list]:
    """Level 3: the reference's 18-field track arrays (accumulate_fm @B35952) rebuilt from fa_track headers and fa_track_point
    rows -- what get_ranked_formants @B35670 returns and P() hands to the callback."""
    out = []
    for t in tracks:
        p = points[int(t["start"]): int(t["start"]) + int(t["len"])]
        bins = [int(x) for x in p["bin"]]
        h = len(bins) - 1                      # points before the last update
        o = bins[-1]
        if h >= 3:
            vel = (o - bins[h - 1] + (bins[h - 2] - bins[h - 1]) + (bins[h - 3] - bins[h - 2])) / 3
        elif h == 2:
            vel = (o - bins[h - 1] + (bins[h - 2] - bins[h - 1])) / 2
        elif h == 1:
            vel = float(o - bins[h - 1])
        else:
            vel = 0.0
        e = [float(x) for x in p["energy"]]
        last = p[-1]
        out.append([float(last["lo"]), float(last["hi"]), float(last["frame"]), float(last["frame"]), float(vel), float(o),
                    float(last["amp"]), [float(x) for x in p["frame"]], [float(x) for x in p["lo"]], [float(x) for x in p["hi"]],
                    [float(b) for b in bins], [float(x) for x in p["amp"]], e, float(sum(e)), float(len(bins)),
                    float(sum(x * b for x, b in zip(e, bins))), 0.0, float(sum(int(a["hi"]) - int(a["lo"]) + 1 for a in p))])
    return out


def segment_callbacks(level: int, step_ms: float, labels: list, res) -> list[tuple]:
    """P() @B28869 on the tables of one utterance: the argument tuples of every callback, in firing order.

    The store index `e` also indexes seg_ci for the time stamps (get_seg_timestamps @B31504 /
    get_syls_timestamps @B31114) -- after a dropped segment the reference pairs store e with seg_ci[e],
    not with the segment that produced it; reproduced here.
    """
    step = step_ms / 1e3
    segs = res.segments
    stored = [s for s in segs if s["stored"] >= 0]
    out = []
    if level == 11:
        # b(0, label, get_clip_timestamps() @B31365, get_utterance_features(u, h) @B107983) after every stored segment;
        # the clip time sums the lengths of ALL seg_ci entries that exist at that moment (dropped ones included)
        k = 0
        for si, s in enumerate(segs):
            if s["stored"] < 0:
                continue
            t = sum(int(x["len"]) for x in segs[: si + 1])
            out.append((0, labels, [int(segs[0]["start"]) * step, (t + 1) * step], list(map(float, res.utterance[k]))))
            k += 1
        return out
    if level == 3:
        # b(e, label, s[e]) when s[e].length > 0: three arguments, no time stamps (P() @B28869, level-3 branch)
        for e, s in enumerate(stored):
            tr = res.syllables[s["first_syllable"]: s["first_syllable"] + s["n_syllables"]]     # fa_track headers at level 3
            if len(tr):
                out.append((e, labels, track_arrays(tr, res.track_points)))
        return out
    for e, s in enumerate(stored):
        ci = segs[e]
        rows = res.formants[s["row_offset"]: s["row_offset"] + s["len"]]
        if level in (13, 12, 10):
            syl = res.syllables[s["first_syllable"]: s["first_syllable"] + s["n_syllables"]]
            if len(syl) == 0:
                continue
            if level == 12:
                # make_coeffs @B34527 returns the rows made before numeric threw (try / catch around the loop); P() fires when
                # there is at least one, with the time stamps of ALL the segment's syllables (j(e) @B31114)
                f0 = int(s["first_syllable"])
                bad = [k for k, y in enumerate(syl) if int(y["reserved"])]
                n_ok = bad[0] if bad else len(syl)
                if n_ok == 0:
                    continue
                payload = [list(map(float, r)) for r in res.features[f0: f0 + n_ok]]
            times = [[to_fixed3((int(ci["start"]) + int(y["start"])) * step), to_fixed3((int(y["len"]) + 1) * step)] for y in syl]
            if level == 12:
                pass
            elif level == 13:
                # feature rows of an utterance are in syllable order
                f0 = int(s["first_syllable"])
                payload = [list(map(float, r)) for r in res.features[f0: f0 + len(syl)]]
            else:
                payload = [[np.array(r, np.float32) for r in rows[int(y["start"]): int(y["start"]) + int(y["len"])]] for y in syl]
            out.append((e, labels, times, payload))
        elif level == 5:
            out.append((e, labels, [int(ci["start"]) * step, (int(ci["len"]) + 1) * step], list(map(float, res.features[e]))))
        elif level == 4:
            if len(rows) == 0:
                continue
            out.append((e, labels, [int(ci["start"]) * step, (int(ci["len"]) + 1) * step], [np.array(r, np.float32) for r in rows]))
    return out


def _engine(cfg: FaConfig, device: int) -> Engine:
    key = (bytes(cfg), device)
    if _state["engine"] is None or _state["engine_key"] != key:
        if _state["engine"] is not None:
            _state["engine"].close()
        _state["engine"] = Engine(cfg, device)
        _state["engine_key"] = key
    return _state["engine"]


def LaunchAudioNodes(context_source, source_obj=None, callback=None, file_labels=None, offline=False, test_play=True,
                     play_offset=None, play_duration=None) -> Future:
    fut: Future = Future()
    file_labels = [] if file_labels is None else file_labels

    def reject(msg: str):
        fut.set_exception(LaunchError(msg))
        return fut

    a = _settings
    if _state["playing"]:
        return reject("Error: Already playing")
    if not (a["N_mel_bins"] or a["N_fft_bins"]):
        return reject("Invalid reset_nodes config")
    bands = a["N_mel_bins"] if a["spec_type"] == 1 else a["N_fft_bins"]
    if not bands:
        return reject("reset_segmentor failed: Invalid spec_bands")
    if not a["output_level"]:
        return reject("Invalid reset_plot config")
    try:
        if context_source == 1 and source_obj is not None:
            pcm, sr = wav.decode_wav(source_obj)
        elif context_source == 4 and source_obj is not None:
            pcm, sr = np.asarray(source_obj["pcm"], np.float32), int(source_obj["sampleRate"])
        else:
            # 2 (<audio> element) and 3 (microphone) need a browser audio graph
            return reject("Invalid audio source")
    except wav.WavError as e:
        return reject(str(e))
    if play_offset or play_duration:
        o = int(round((play_offset or 0) * sr))
        n = pcm.size - o if not play_duration else int(round(play_duration * sr))
        pcm = pcm[o: o + max(n, 0)]
    _state["playing"] = True
    _state["labels"] = []
    try:
        cfg = _fa_config()
        eng = _engine(cfg, int(a["device"]))
        eng.reset()
        eng.submit(0, pcm, sr)
        eng.run()
        eng.sync()
        level = int(a["output_level"])
        if level <= 2:
            if not test_play and callback:
                sp = eng.spectrum(0)
                callback(0, file_labels, [0.0, sp.shape[0] * a["window_step"] / 1e3], sp)
        else:
            res = eng.result(0)
            calls = segment_callbacks(level, float(a["window_step"]), file_labels, res)
            _state["labels"] = [list(file_labels) for _ in calls]
            if not test_play and callback:
                for args in calls:
                    if _state["stop"] is not None:
                        break
                    callback(*args)
        fut.set_result(True)
    except Exception as e:  # noqa: BLE001  (the reference rejects with the error)
        fut.set_exception(LaunchError(str(e)))
    finally:
        _state["playing"] = False
        _state["stop"] = None
    return fut


class LiveSession:
    """Incremental callbacks for audio that is still arriving (the reference's microphone / <audio> sources: P() @B28869 fires
    after every pause while the stream runs).

    The segmentor is causal, so the segments a PREFIX of the stream finalises on its own -- without segment_truncate @B30800,
    which the reference only runs when the source stops -- are exactly the ones the reference has called back by then.  push()
    appends a chunk, re-analyses the stream so far with truncation off (at > 300 000 x real time: 0.2 ms per minute of audio)
    and fires the callbacks that are new, with the same arguments a one-shot analysis of the whole stream would give them;
    stop() runs once more with truncation on and fires the rest."""

    def __init__(self, sample_rate: int, callback, file_labels=None, test_play: bool = False):
        self.sr = int(sample_rate)
        self.callback = callback
        self.labels = [] if file_labels is None else file_labels
        self.test_play = test_play
        self.cfg = _fa_config()
        self.level = int(_settings["output_level"])
        self.step = float(_settings["window_step"])
        if self.level <= 2:
            raise LaunchError("LiveSession delivers segment callbacks: output_level must be >= 3")
        self.eng = _engine(self.cfg, int(_settings["device"]))
        self.chunks: list[np.ndarray] = []
        self.fired = 0
        self.stopped = False

    def _run(self, truncate: bool) -> int:
        pcm = np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.float32)
        eng = self.eng
        eng.reset()
        eng.set_truncate(truncate)
        try:
            eng.submit(0, pcm, self.sr)
            eng.run()
            eng.sync()
            calls = segment_callbacks(self.level, self.step, self.labels, eng.result(0))
        finally:
            eng.set_truncate(True)
        new = calls[self.fired:]
        self.fired = len(calls)
        if not self.test_play and self.callback:
            for args in new:
                self.callback(*args)
        return len(new)

    def push(self, pcm_chunk) -> int:
        """Append float32 mono samples; returns how many callbacks fired."""
        if self.stopped:
            raise LaunchError("the session was stopped")
        self.chunks.append(np.ascontiguousarray(pcm_chunk, np.float32))
        return self._run(False)

    def stop(self) -> int:
        """The source stopped (disconnect_nodes @B21559 -> segment_truncate): fires what the truncation finalises."""
        self.stopped = True
        return self._run(True)


def StopAudioNodes(reason: str = "no reason") -> None:
    """disconnect_nodes @B21559: only has an effect while playing."""
    if _state["playing"]:
        _state["stop"] = reason


def set_predicted_label_for_segment(seg_index: int, label_index: int, predicted_label) -> None:
    """set_segments_label (module 3 `G`): pads the segment's label list with -1 up to label_index, then sets it."""
    lab = _state["labels"][seg_index]
    while len(lab) < label_index:
        lab.append(-1)
    if len(lab) == label_index:
        lab.append(predicted_label)
    else:
        lab[label_index] = predicted_label


def get_segment_labels(seg_index: int):
    return _state["labels"][seg_index]
