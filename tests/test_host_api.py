"""Host-side mirror of the formantanalyzer API: configure() truthiness rules, string rejections, WAV decoding,
toFixed(3), and the callback shapes / order of P() checked against the literal transliteration."""
import json
import numpy as np
import pytest

from oracle import oracle
from oracle.literal.refmodules import Segmentor
from webspeechanalyzer_b200 import FaConfig, api, synth_speech, wav
from webspeechanalyzer_b200.engine import UtteranceResult


@pytest.fixture(autouse=True)
def _fresh():
    api.reset_defaults()
    yield
    api.reset_defaults()


def test_configure_truthiness_rules():
    api.configure({"output_level": 0, "window_step": 0, "f_min": 0, "high_f_emph": 0, "auto_noise_gate": False,
                   "voiced_min_dB": 0, "spec_type": 3})
    s = api._settings
    assert s["output_level"] == 4 and s["window_step"] == 25          # falsy values cannot be set (@B3292)
    assert s["f_min"] == 0 and s["auto_noise_gate"] is False and s["voiced_min_dB"] == 0 and s["spec_type"] == 3
    api.configure({"output_level": 13, "window_step": 15, "fftSize": 1024, "smoothingTimeConstant": 0})
    c = api._fa_config()
    assert (c.output_level, c.window_step_ms, c.fft_size, c.smoothing, c.spec_type, c.bands) == (13, 15.0, 1024, 0.0, 3, 256)


def test_rejections_are_the_reference_strings():
    assert str(api.LaunchAudioNodes(2, object()).exception()) == "Invalid audio source"
    assert str(api.LaunchAudioNodes(3).exception()) == "Invalid audio source"
    assert str(api.LaunchAudioNodes(1, None).exception()) == "Invalid audio source"
    assert str(api.LaunchAudioNodes(1, b"not a wav").exception()) == "Unable to decode audio data"
    api._state["playing"] = True
    assert str(api.LaunchAudioNodes(4, {"pcm": np.zeros(10, np.float32), "sampleRate": 16000}).exception()) == "Error: Already playing"
    api._state["playing"] = False
    api.StopAudioNodes("not playing: no effect")
    assert api._state["stop"] is None


def test_wav_roundtrip_and_formats():
    p = synth_speech(8000, 16000, 1, 0)
    x, sr = wav.decode_wav(wav.encode_wav_pcm16(p, 16000))
    assert sr == 16000 and x.dtype == np.float32 and np.abs(x - p).max() <= 0.5 / 32768 + 1e-7
    import struct
    raw = p.astype("<f4").tobytes()
    f32 = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32) \
        + b"data" + struct.pack("<I", len(raw)) + raw
    y, _ = wav.decode_wav(f32)
    assert np.array_equal(y, p)
    st = np.stack([p, -p], axis=1).astype("<f4").tobytes()
    f32s = b"RIFF" + struct.pack("<I", 36 + len(st)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 3, 2, 16000, 128000, 8, 32) \
        + b"data" + struct.pack("<I", len(st)) + st
    z, _ = wav.decode_wav(f32s)
    assert np.abs(z).max() == 0.0


def test_to_fixed3_matches_js():
    assert api.to_fixed3(0.0625) == "0.063" and api.to_fixed3(1.305) == "1.305" and api.to_fixed3(0.015 * 87) == "1.305"


def as_result(an) -> UtteranceResult:
    return UtteranceResult({}, an.segments, an.formants, an.energy, an.syllables, an.features, getattr(an, "utterance", None))


@pytest.mark.parametrize("level,step", [(13, 15.0), (5, 25.0), (4, 25.0), (10, 15.0)])
def test_callback_shapes_and_order_equal_literal_P(level, step):
    sr = 16000
    cfg = FaConfig.default(output_level=level, window_step_ms=step)
    pcm = np.concatenate([synth_speech(4 * sr, sr, 21, u) for u in range(3)])
    fe, an = oracle.analyze_pcm(cfg, pcm, sr)
    S = Segmentor(level, cfg.bands, 200, step, 200, 50, True, 100, 10, None, True, ["lab"])
    for f in fe["frames"]:
        S.spectrum_push(f)
    S.segment_truncate()
    calls = api.segment_callbacks(level, step, ["lab"], as_result(an))
    assert len(calls) == len(S.events) > 0
    for mine, ref in zip(calls, S.events):
        assert mine[0] == ref[0] and mine[1] == ref[1]
        if level in (13, 10):
            assert mine[2] == ref[2]                                   # toFixed(3) strings
            assert all(isinstance(t, str) for pair in mine[2] for t in pair)
        else:
            assert mine[2] == ref[2]                                   # numbers: start*step, (len+1)*step
        if level == 13:
            assert np.array_equal(np.array(mine[3]), np.array(ref[3]), equal_nan=True) and len(mine[3][0]) == 53
        elif level == 5:
            assert np.array_equal(np.array(mine[3]), np.array(ref[3]), equal_nan=True) and len(mine[3]) == 53
        elif level == 4:
            assert np.array_equal(np.stack(mine[3]), np.stack(ref[3])) and mine[3][0].dtype == np.float32
        elif level == 10:
            assert len(mine[3]) == len(ref[3])
            for a, b in zip(mine[3], ref[3]):
                assert np.array_equal(np.stack(a), np.stack(b))


def test_dropped_segment_misaligns_times_like_the_reference():
    """Quirk 15: straighten_formants throws when a track point's frame index >= len; seg_ci keeps the entry but the
    stores do not, so later callbacks pair store e with seg_ci[e]."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    from framegen import dropped_segment_frames
    B = 128
    frames = list(dropped_segment_frames(B))
    cfg = FaConfig.default(output_level=5)
    an = oracle.analyze_frames(cfg, np.stack(frames))
    S = Segmentor(5, B, 200, 25.0, 200, 50, True, 100, 10, None, True, [])
    for f in frames:
        S.spectrum_push(f)
    S.segment_truncate()
    assert [tuple(x) for x in S.u] == an.seg_ci
    stored = [int(s["stored"]) for s in an.segments]
    assert stored == [-1, 0] and an.seg_ci == [(8, 3), (22, 12)]   # the drop happened
    calls = api.segment_callbacks(5, 25.0, [], as_result(an))
    assert len(calls) == len(S.events) == 1
    # store 0 belongs to seg_ci[1] but is stamped with seg_ci[0]'s time, exactly like the reference
    assert calls[0][2] == S.events[0][2] == [8 * 0.025, (3 + 1) * 0.025]
    assert calls[0][3][0] == 12.0


# ------------------------------------------------------------------ export in the web app's storage format
def test_export_rows_json_csv_follow_localstore():
    from webspeechanalyzer_b200 import export
    assert export.js_number(101.0) == "101" and export.js_number(0.1 + 0.2) == "0.30000000000000004"
    assert export.js_number(1.5e-7) == "1.5e-7" and export.js_number(1e21) == "1e+21" and export.js_number(float("nan")) == "NaN"
    sr = 16000
    pcm = synth_speech(5 * sr, sr, 1, 0)
    # level 13: one stored row per syllable, key "<db>#<file>#<si + ph/100>" (src/index.js:53, src/localstore.js:41-42)
    cfg = FaConfig.default(output_level=13, window_step_ms=15.0)
    _, an = oracle.analyze_pcm(cfg, pcm, sr)
    calls = api.segment_callbacks(13, 15.0, ["f.wav"], as_result(an))
    rows = export.stored_rows(13, 7, "f.wav", calls)
    assert len(rows) == an.features.shape[0] > 0 and rows[0]["key"] == "7#f.wav#0" and all(len(r["features"]) == 53 for r in rows)
    multi = [c for c in calls if len(c[3]) > 1]
    if multi:
        assert f"7#f.wav#{export.js_number(multi[0][0] + 0.01)}" in [r["key"] for r in rows]
    import json
    doc = json.loads(export.to_json(rows, origin=["f.wav"]))
    assert set(doc[0]) == {"file", "seg", "time", "features", "origin", "true", "pred"}        # localstore.js:883
    assert doc[0]["file"] == "f.wav" and doc[0]["seg"] == "0" and doc[0]["time"] == calls[0][2][0]  # toFixed(3) strings
    got = np.array([[np.nan if v is None else v for v in d["features"]] for d in doc])
    assert np.array_equal(got, an.features, equal_nan=True)                                   # repr round-trips the doubles
    csv = export.to_csv(rows).split("\r\n")
    assert csv[0] == "file,seg,t0,td," + "".join(f"x{i}," for i in range(53))                  # localstore.js:900, 934
    assert csv[1].startswith(f"f.wav,0,{calls[0][2][0][0]},{calls[0][2][0][1]},") and csv[1].count(",") == 4 + 53
    # level 5: one row per segment, numeric time stamps
    cfg = FaConfig.default(output_level=5)
    _, an = oracle.analyze_pcm(cfg, pcm, sr)
    rows = export.stored_rows(5, "db", "f.wav", api.segment_callbacks(5, 25.0, [], as_result(an)))
    assert [r["seg"] for r in rows] == [str(i) for i in range(len(rows))] and len(rows) == an.features.shape[0]
    assert export.to_csv(rows).split("\r\n")[1].split(",")[2] == export.js_number(an.seg_ci[0][0] * 0.025)
    # level 10: float32 mean of the syllable's rows (src/index.js:74-86); level 11: the last cumulative row, segment 0
    cfg = FaConfig.default(output_level=10, window_step_ms=15.0)
    _, an = oracle.analyze_pcm(cfg, pcm, sr)
    calls = api.segment_callbacks(10, 15.0, [], as_result(an))
    rows = export.stored_rows(10, 1, "f.wav", calls)
    fr = np.stack(calls[0][3][0]).astype(np.float32)
    acc = fr[0].copy()
    for r in fr[1:]:
        acc += r
    assert len(rows[0]["features"]) == 9 and np.array_equal(np.array(rows[0]["features"], np.float32),
                                                            np.where(acc != 0, acc / np.float32(len(fr)), acc))
    cfg = FaConfig.default(output_level=11, window_step_ms=15.0)
    _, an = oracle.analyze_pcm(cfg, pcm, sr)
    rows = export.stored_rows(11, 1, "f.wav", api.segment_callbacks(11, 15.0, [], as_result(an)))
    assert len(rows) == 1 and rows[0]["seg"] == "0" and np.array_equal(rows[0]["features"], an.utterance[-1])
    # level 4 is not stored by the app; rows of the wrong width are refused
    assert export.stored_rows(4, 1, "f.wav", [(0, [], [0.0, 1.0], [np.zeros(9, np.float32)])]) == []
    assert export.stored_rows(5, 1, "f.wav", [(0, [], [0.0, 1.0], [1.0] * 52)]) == []


# ---------------------------------------------------------------------------- the Node shim's JavaScript, executed
def _norm(x):
    """callback payloads -> nested lists of floats / strings (numpy rows, tuples and JS arrays alike)."""
    import numpy as _np
    if isinstance(x, _np.ndarray):
        return [_norm(v) for v in x.tolist()]
    if isinstance(x, (list, tuple)):
        return [_norm(v) for v in x]
    if isinstance(x, (int, float, _np.integer, _np.floating)):
        return float(x)
    return x


@pytest.mark.parametrize("level", [3, 4, 5, 10, 11, 12, 13])
def test_node_shim_segment_callbacks_executed_by_minijs(level):
    """webspeechanalyzer_b200/node/index.js cannot run here (no Node), but its segmentCallbacks() -- the function that turns
    the addon's tables into the reference's callback arguments at every level -- is plain JavaScript: oracle/minijs executes it
    on the oracle's tables and the result must equal the Python twin's (which is pinned to the reference's own P())."""
    import os
    from oracle import oracle
    from oracle.minijs.interp import Interp, JSArray, JSObject, JSTyped, UNDEF
    from oracle.minijs.run_reference import _to_py
    from webspeechanalyzer_b200 import FaConfig, synth_speech
    from webspeechanalyzer_b200.engine import UtteranceResult
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "webspeechanalyzer_b200", "node", "index.js")).read()
    fn = src[src.index("function segmentCallbacks"): src.index("function LaunchAudioNodes")]
    sr, step = 16000, 15.0
    pcm = np.concatenate([synth_speech(4 * sr, sr, 5, k) for k in range(2)])
    cfg = FaConfig.default(output_level=level, window_step_ms=step)
    an = oracle.analyze_frames(cfg, oracle.frontend(cfg, pcm, sr, spectrum=False)["frames"])
    res = UtteranceResult({}, an.segments, an.formants, an.energy, an.syllables, an.features, an.utterance, an.track_points)
    labels = [3.0, 1.0]
    want = api.segment_callbacks(level, step, labels, res)
    assert len(want) >= 2
    it = Interp()
    it.run("var a={window_step:%r};" % step + fn)
    segs = JSArray([JSObject({"start": float(s["start"]), "len": float(s["len"]), "stored": float(s["stored"]),
                              "nSyllables": float(s["n_syllables"]), "firstSyllable": float(s["first_syllable"]),
                              "rowOffset": float(s["row_offset"])}) for s in an.segments])
    syls = JSArray([JSObject({"storedSeg": float(y["stored_seg"]), "start": float(y["start"]), "len": float(y["len"]),
                              "flag": float(y["reserved"])}) for y in an.syllables])
    feats = an.utterance if level == 11 else an.features
    props = {"segments": segs, "syllables": syls, "formants": JSTyped("Float32Array", [float(x) for x in an.formants.ravel()]),
             "features": JSTyped("Float64Array", [float(x) for x in feats.ravel()])}
    if level == 3:
        tp = an.track_points
        props["trackPoints"] = JSTyped("Float64Array", [float(x) for c in ("frame", "lo", "hi", "bin", "amp", "energy") for x in tp[c]])
    got = _to_py(it.call(it.globals.vars["segmentCallbacks"], UNDEF, [float(level), JSObject(props), JSArray(list(labels))]))
    a, b = _norm(got), _norm(want)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(np.array(x[0]), np.array(y[0])) and x[1] == y[1]
        assert json.dumps(x[2:], allow_nan=True) == json.dumps(y[2:], allow_nan=True)


def test_node_shim_configure_executed_by_minijs_matches_the_reference_module():
    """node/index.js configure() run by oracle/minijs on the settings objects of tests/golden/ref_js_api.json: it must leave
    what the reference's own configure() (@B3292, executed the same way) left -- same rule as the Python mirror's test: for
    partial objects the reference stores `undefined` into the `null !== x` fields, the shims keep the previous value."""
    import os
    from conftest import GOLDEN
    from oracle.minijs.interp import Interp, JSObject, UNDEF
    from oracle.minijs.run_reference import _to_py
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "webspeechanalyzer_b200", "node", "index.js")).read()
    defaults = src[src.index("const DEFAULTS = Object.freeze({"): src.index("let a = Object.assign")]
    defaults = defaults.replace("const DEFAULTS = Object.freeze({", "var a = {").replace("});", "};")
    fn = src[src.index("function configure(e)"): src.index("function decodeWav")]
    doc = json.load(open(os.path.join(GOLDEN, "ref_js_api.json")))
    fields = ("plot_enable", "spec_type", "output_level", "plot_len", "f_min", "f_max", "N_fft_bins", "N_mel_bins", "window_width",
              "window_step", "pause_length", "min_seg_length", "auto_noise_gate", "voiced_max_dB", "voiced_min_dB", "pre_norm_gain",
              "high_f_emph")
    nullable = ("spec_type", "f_min", "high_f_emph", "auto_noise_gate", "voiced_min_dB")

    def js(v):
        return float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else (None if v is None else v)

    for case in doc["configure"]:
        it = Interp()
        it.run(defaults + fn)
        before = _to_py(it.globals.vars["a"])
        it.call(it.globals.vars["configure"], UNDEF, [JSObject({k: js(v) for k, v in case["cfg"].items()})])
        mine, ref = _to_py(it.globals.vars["a"]), case["settings"]
        partial = not all(f in case["cfg"] for f in nullable)
        for f in fields:
            if partial and ref[f] is None and f in nullable:
                assert mine[f] == before[f]
                continue
            assert mine[f] == ref[f], (f, mine[f], ref[f])


# ---------------------------------------------------------------------------- incremental callbacks (live streams)
@pytest.mark.parametrize("level", [3, 5, 11, 13])
def test_prefix_without_truncate_gives_the_callbacks_fired_so_far(level):
    """The premise of api.LiveSession / fa_set_truncate: the segmentor is causal, so what a PREFIX of a stream finalises without
    segment_truncate is a prefix of what the whole stream calls back -- same arguments, same order (oracle on both sides)."""
    sr, step = 16000, 25.0
    pcm = np.concatenate([synth_speech(4 * sr, sr, 7, k) for k in range(4)])
    cfg = FaConfig.default(output_level=level, window_step_ms=step)
    frames = oracle.frontend(cfg, pcm, sr, spectrum=False)["frames"]

    def calls(fr, truncate):
        an = oracle.analyze_frames(cfg, fr, truncate=truncate)
        res = UtteranceResult({}, an.segments, an.formants, an.energy, an.syllables, an.features, an.utterance, an.track_points)
        return [_norm(c) for c in api.segment_callbacks(level, step, [1.0], res)]

    full = calls(frames, True)
    assert len(full) >= 3
    seen = 0
    for n in range(0, frames.shape[0] + 1, 13):
        part = calls(frames[:n], False)
        assert json.dumps(part) == json.dumps(full[: len(part)])      # a prefix of the final list, bit for bit
        assert len(part) >= seen                                       # and it only grows
        seen = len(part)
    assert seen >= len(full) - 1                                       # only the truncated tail waits for the source to stop
    assert 0 < len(calls(frames[: frames.shape[0] // 2], False)) < len(full)
