"""
TEST INFRASTRUCTURE ONLY.  ctypes front of the CPU oracle (oracle/fa_oracle.c).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from webspeechanalyzer_b200._ctypes_defs import FaConfig, FaSegment, FaSyllable, N_FEATURES

from . import build

_lib = C.CDLL(build.ensure_built())
_cfgp = C.POINTER(FaConfig)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)

_lib.fao_hop.argtypes = [_cfgp, C.c_int]
_lib.fao_bands.argtypes = [_cfgp]
_lib.fao_num_frames.argtypes = [_cfgp, C.c_int, C.c_int64]
_lib.fao_frontend.argtypes = [_cfgp, _f32p, C.c_int64, C.c_int, _f32p, _f32p, _u32p]
_lib.fao_frontend_f64.argtypes = [_cfgp, _f32p, C.c_int64, C.c_int, _f64p]
_lib.fao_analyze_frames.argtypes = [_cfgp, _u32p, C.c_int, C.c_int]
_lib.fao_analyze_frames.restype = C.c_void_p
_lib.fao_free.argtypes = [C.c_void_p]
_lib.fao_counts.argtypes = [C.c_void_p, _i32p]
_lib.fao_get_segments.argtypes = [C.c_void_p, C.POINTER(FaSegment)]
_lib.fao_track_stats.argtypes = [C.c_void_p, _i32p]
_lib.fao_get_formants.argtypes = [C.c_void_p, _f32p, _f32p]
_lib.fao_get_syllables.argtypes = [C.c_void_p, C.POINTER(FaSyllable)]
_lib.fao_get_features.argtypes = [C.c_void_p, _f64p]
_lib.fao_track_counts.argtypes = [C.c_void_p, _i32p]
_lib.fao_get_tracks.argtypes = [C.c_void_p, C.c_void_p]
_lib.fao_get_track_points.argtypes = [C.c_void_p, C.c_void_p]
_lib.fao_get_callbacks.argtypes = [C.c_void_p, _i32p]
_lib.fao_utterance_rows.argtypes = [C.c_void_p]
_lib.fao_get_utterance_features.argtypes = [C.c_void_p, _f64p]
_lib.fao_get_trace.argtypes = [C.c_void_p, _i32p, _i32p, _f64p, _f64p, _f64p, _i32p, _i32p, _i32p]
_lib.fao_peak_candidates.argtypes = [_u32p, C.c_int, _u32p, _f64p]
_lib.fao_run_batch.argtypes = [_cfgp, _f32p, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
_lib.fao_run_batch.restype = C.c_int64


def _p(a, t):
    return a.ctypes.data_as(t)


def hop(cfg: FaConfig, sr: int) -> int:
    return _lib.fao_hop(C.byref(cfg), sr)


def num_frames(cfg: FaConfig, sr: int, n: int) -> int:
    return _lib.fao_num_frames(C.byref(cfg), sr, n)


def frontend(cfg: FaConfig, pcm: np.ndarray, sr: int, spectrum=True, smooth=False, frames=True):
    """Canonical float32 front end.  Returns dict(spectrum=[F,N/2] dB, smooth=[F,N/2], frames=[F,B] uint32)."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    F = num_frames(cfg, sr, pcm.size)
    M = cfg.fft_size // 2
    out = {}
    sp = np.empty((F, M), np.float32) if spectrum else None
    sm = np.empty((F, M), np.float32) if smooth else None
    fr = np.empty((F, cfg.bands), np.uint32) if frames else None
    rc = _lib.fao_frontend(C.byref(cfg), _p(pcm, _f32p), pcm.size, sr, _p(sp, _f32p) if spectrum else None,
                           _p(sm, _f32p) if smooth else None, _p(fr, _u32p) if frames else None)
    if rc < 0:
        raise ValueError("fao_frontend failed (bad fft_size?)")
    out["spectrum"], out["smooth"], out["frames"] = sp, sm, fr
    return out


def byte_view(spec_db_unclamped: np.ndarray, cfg: FaConfig) -> np.ndarray:
    """AnalyserNode.getByteFrequencyData of the (unclamped) float dB rows: trunc(clamp(255 / (max - min) * (Y - min), 0, 255)),
    float32 arithmetic, -inf -> 0 (DESIGN.md front-end spec)."""
    y = np.asarray(spec_db_unclamped, np.float32)
    scale = np.float32(255.0 / (cfg.max_db - cfg.min_db))
    with np.errstate(invalid="ignore"):
        t = (y - np.float32(cfg.min_db)) * scale
    t = np.where(np.isnan(t), np.float32(0), t)
    return np.minimum(np.maximum(t, np.float32(0)), np.float32(255)).astype(np.uint8)


def frontend_f64(cfg: FaConfig, pcm: np.ndarray, sr: int) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, np.float32)
    F = num_frames(cfg, sr, pcm.size)
    sp = np.empty((F, cfg.fft_size // 2), np.float64)
    if _lib.fao_frontend_f64(C.byref(cfg), _p(pcm, _f32p), pcm.size, sr, _p(sp, _f64p)) < 0:
        raise ValueError("fao_frontend_f64 failed")
    return sp


def peak_candidates(frame: np.ndarray):
    frame = np.ascontiguousarray(frame, np.uint32)
    packed = np.zeros(frame.size, np.uint32)
    g = C.c_double(0)
    n = _lib.fao_peak_candidates(_p(frame, _u32p), frame.size, _p(packed, _u32p), C.byref(g))
    return packed[:n].copy(), g.value


@dataclass
class Analysis:
    frames: int
    bands: int
    segments: np.ndarray            # structured FaSegment array (seg_ci order)
    formants: np.ndarray            # [rows, 9] float32
    energy: np.ndarray              # [rows, 3] float32
    syllables: np.ndarray           # structured FaSyllable array
    features: np.ndarray            # [rows, 53] float64
    callbacks: np.ndarray           # store indices in firing order
    trace: dict = field(default_factory=dict)
    utterance: np.ndarray = None    # [callbacks, 264] float64 (level 11)
    max_live_tracks: int = 0        # most tracks still matchable after one accumulate_fm call (GPU slot demand)
    max_peaks: int = 0              # most accepted peaks in one accumulate_fm call
    track_points: np.ndarray = None  # level 3: TRACK_POINT_DTYPE rows; `syllables` then holds the fa_track headers

    @property
    def seg_ci(self):
        return [(int(s["start"]), int(s["len"])) for s in self.segments]


SEG_DTYPE = np.dtype([("start", "<i4"), ("len", "<i4"), ("stored", "<i4"), ("n_syllables", "<i4"),
                      ("first_syllable", "<i4"), ("row_offset", "<i4"), ("ymax", "<f8"), ("vmin", "<f8"),
                      ("cs_ratio", "<f8")])
SYL_DTYPE = np.dtype([("stored_seg", "<i4"), ("start", "<i4"), ("len", "<i4"), ("reserved", "<i4")])
TRACK_POINT_DTYPE = np.dtype([("frame", "<i4"), ("lo", "<i2"), ("hi", "<i2"), ("bin", "<i2"), ("reserved", "<i2"), ("amp", "<u4"),
                              ("energy", "<f8")], align=True)
assert SEG_DTYPE.itemsize == C.sizeof(FaSegment) and SYL_DTYPE.itemsize == C.sizeof(FaSyllable) and TRACK_POINT_DTYPE.itemsize == 24


def analyze_frames(cfg: FaConfig, frames: np.ndarray, trace: bool = False, truncate: bool = True) -> Analysis:
    """truncate=False: the frames are the prefix of a stream that is still running (no segment_truncate @B30800 at the end)."""
    frames = np.ascontiguousarray(frames, np.uint32)
    F = frames.shape[0]
    assert F == 0 or frames.shape[1] == cfg.bands
    R = _lib.fao_analyze_frames(C.byref(cfg), _p(frames, _u32p), F, (1 if trace else 0) | (0 if truncate else 2))
    try:
        cnt = (C.c_int * 8)()
        _lib.fao_counts(R, cnt)
        _, nseg, _nst, nrows, nsyl, nfeat, ncb, B = list(cnt)
        segs = np.zeros(nseg, SEG_DTYPE)
        if nseg:
            _lib.fao_get_segments(R, segs.ctypes.data_as(C.POINTER(FaSegment)))
        Fm = np.zeros((nrows, 9), np.float32)
        Eg = np.zeros((nrows, 3), np.float32)
        if nrows:
            _lib.fao_get_formants(R, _p(Fm, _f32p), _p(Eg, _f32p))
        syl = np.zeros(nsyl, SYL_DTYPE)
        if nsyl:
            _lib.fao_get_syllables(R, syl.ctypes.data_as(C.POINTER(FaSyllable)))
        feat = np.zeros((nfeat, 23 if cfg.output_level == 12 else N_FEATURES), np.float64)
        if nfeat:
            _lib.fao_get_features(R, _p(feat, _f64p))
        cb = np.zeros(ncb, np.int32)
        if ncb:
            _lib.fao_get_callbacks(R, _p(cb, _i32p))
        tr = {}
        if trace and F:
            tr = {k: np.zeros(F, np.int32) for k in ("n", "p", "cstart", "cci", "nofm")}
            tr.update({k: np.zeros(F, np.float64) for k in ("h", "v", "y")})
            _lib.fao_get_trace(R, _p(tr["n"], _i32p), _p(tr["p"], _i32p), _p(tr["h"], _f64p), _p(tr["v"], _f64p),
                               _p(tr["y"], _f64p), _p(tr["cstart"], _i32p), _p(tr["cci"], _i32p), _p(tr["nofm"], _i32p))
        nu = _lib.fao_utterance_rows(R)
        utt = np.zeros((nu, 264), np.float64)
        if nu:
            _lib.fao_get_utterance_features(R, _p(utt, _f64p))
        ts = (C.c_int * 2)()
        _lib.fao_track_stats(R, ts)
        pts = None
        if cfg.output_level == 3:      # the ranked tracks (headers in the fa_syllable layout) and their points
            tc = (C.c_int * 2)()
            _lib.fao_track_counts(R, tc)
            syl = np.zeros(tc[0], SYL_DTYPE)
            pts = np.zeros(tc[1], TRACK_POINT_DTYPE)
            if tc[0]:
                _lib.fao_get_tracks(R, syl.ctypes.data)
                _lib.fao_get_track_points(R, pts.ctypes.data)
        return Analysis(F, B, segs, Fm, Eg, syl, feat, cb, tr, utt, int(ts[0]), int(ts[1]), pts)
    finally:
        _lib.fao_free(R)


def analyze_pcm(cfg: FaConfig, pcm: np.ndarray, sr: int, trace: bool = False):
    fe = frontend(cfg, pcm, sr, spectrum=bool(cfg.want_spectrum) or cfg.output_level <= 2)
    return fe, analyze_frames(cfg, fe["frames"], trace=trace)


def run_batch(cfg: FaConfig, pcm: np.ndarray, offsets: np.ndarray, sr: int, threads: int) -> int:
    """Whole path over a batch with OpenMP over utterances; returns total frames (bench CPU baseline)."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    return int(_lib.fao_run_batch(C.byref(cfg), _p(pcm, _f32p), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                  offsets.size - 1, sr, threads, None))
