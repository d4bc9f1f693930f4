"""Batch engine: a thin object wrapper over the C-ABI handle (one handle per GPU, one batch at a time)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import FaError
from ._ctypes_defs import FaConfig, FaCounts, N_CURVE_FEATURES, N_FEATURES, N_UTT_FEATURES

SEG_DTYPE = np.dtype([("start", "<i4"), ("len", "<i4"), ("stored", "<i4"), ("n_syllables", "<i4"),
                      ("first_syllable", "<i4"), ("row_offset", "<i4"), ("ymax", "<f8"), ("vmin", "<f8"),
                      ("cs_ratio", "<f8")])
SYL_DTYPE = np.dtype([("stored_seg", "<i4"), ("start", "<i4"), ("len", "<i4"), ("reserved", "<i4")])
TRACK_POINT_DTYPE = np.dtype([("frame", "<i4"), ("lo", "<i2"), ("hi", "<i2"), ("bin", "<i2"), ("reserved", "<i2"), ("amp", "<u4"),
                              ("energy", "<f8")], align=True)     # fa_track_point (include/fa_b200.h)
COUNTS_DTYPE = np.dtype([("samples", "<i8"), ("sample_rate", "<i4"), ("hop", "<i4"), ("frames", "<i4"), ("bands", "<i4"),
                         ("segments", "<i4"), ("stored_segments", "<i4"), ("formant_rows", "<i4"), ("syllables", "<i4"),
                         ("feature_rows", "<i4"), ("overflow", "<i4")])


@dataclass
class UtteranceResult:
    counts: dict
    segments: np.ndarray      # SEG_DTYPE, seg_ci order
    formants: np.ndarray      # [rows, 9] float32
    energy: np.ndarray        # [rows, 3] float32
    syllables: np.ndarray     # SYL_DTYPE
    features: np.ndarray      # [rows, 53] float64
    utterance: np.ndarray | None = None   # [stored segments, 264] float64, cumulative (level 11)
    track_points: np.ndarray | None = None   # level 3: TRACK_POINT_DTYPE rows; `syllables` then holds the fa_track headers

    @property
    def seg_ci(self):
        return [(int(s["start"]), int(s["len"])) for s in self.segments]


def synth_speech(n_samples: int, sample_rate: int, seed: int, utt_index: int) -> np.ndarray:
    """Synthetic glottal-pulse speech (csrc/fa_synth.cpp); host code, needs no GPU."""
    out = np.empty(n_samples, np.float32)
    rc = _capi.synth_lib().fa_synth_speech(out.ctypes.data, n_samples, sample_rate, seed, utt_index)
    if rc != 0:
        raise FaError(rc, "fa_synth_speech")
    return out


def synth_speech_i16_batch(dst: np.ndarray, n_utt: int, n_samples: int, sample_rate: int, seed: int, first_index: int,
                           index_stride: int = 1, threads: int = 1) -> np.ndarray:
    """n_utt utterances of 16-bit PCM written back to back into `dst` (int16, at least n_utt * n_samples; may be page-locked
    caller memory); utterance i carries index first_index + i * index_stride.  OpenMP over utterances, host code."""
    assert dst.dtype == np.int16 and dst.flags.c_contiguous and dst.size >= n_utt * n_samples
    rc = _capi.synth_lib().fa_synth_speech_i16_batch(dst.ctypes.data, n_utt, n_samples, sample_rate, seed, first_index,
                                                     index_stride, threads)
    if rc != 0:
        raise FaError(rc, "fa_synth_speech_i16_batch")
    return dst


class PinnedBuffer:
    """Page-locked host memory from the C-ABI (fa_host_alloc), as a numpy array; write_combined for H2D-only buffers."""

    def __init__(self, shape, dtype, write_combined: bool = False):
        self._lib = _capi.lib()
        dt = np.dtype(dtype)
        n = int(np.prod(shape))
        self._p = self._lib.fa_host_alloc(max(1, n * dt.itemsize), 1 if write_combined else 0)
        if not self._p:
            raise MemoryError("fa_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_byte * (n * dt.itemsize)).from_address(self._p)).view(dt).reshape(shape)

    def free(self):
        if getattr(self, "_p", None):
            self.array = None
            self._lib.fa_host_free(C.c_void_p(self._p))
            self._p = None

    __del__ = free


class Engine:
    """One GPU, one stream.  submit() utterances, run(), then fetch per-utterance results."""

    def __init__(self, cfg: FaConfig, device: int = 0):
        self._lib = _capi.lib()
        self.cfg = cfg.copy()
        self._h = C.c_void_p()
        rc = self._lib.fa_create(C.byref(self.cfg), device, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise FaError(rc, self._lib.fa_status_string(rc).decode())
        self.device = device
        self._keep = []      # zero-copy buffers of the batch in flight (page-locked caller memory must outlive fa_sync)
        self._sink = None

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.fa_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise FaError(rc, (self._lib.fa_last_error(self._h) or b"").decode() or
                          self._lib.fa_status_string(rc).decode())
        return rc

    # -- batch --
    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self._lib.fa_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_d2h_stream(self, cuda_stream_ptr: int | None):
        """Stream of the spectrum-sink D2H copies; engines sharing one send their batches back one after the other."""
        self._check(self._lib.fa_set_d2h_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def reset(self):
        self._check(self._lib.fa_reset(self._h))   # waits for the stream: nothing reads the old batch's buffers any more
        self._keep.clear()

    def submit(self, utt_id: int, pcm: np.ndarray, sample_rate: int) -> int:
        if pcm.dtype == np.int16:
            pcm = np.ascontiguousarray(pcm)
            return self._check(self._lib.fa_submit_pcm_i16(self._h, utt_id, pcm.ctypes.data, pcm.size, sample_rate))
        pcm = np.ascontiguousarray(pcm, np.float32)
        return self._check(self._lib.fa_submit_pcm(self._h, utt_id, pcm.ctypes.data, pcm.size, sample_rate))

    def submit_frames(self, utt_id: int, frames: np.ndarray) -> int:
        """The segmentor's own input (spectrum_push @B30392): [n_frames, bands] uint32; the spectrum stage is skipped."""
        frames = np.ascontiguousarray(frames, np.uint32)
        assert frames.ndim == 2
        return self._check(self._lib.fa_submit_frames(self._h, utt_id, frames.ctypes.data, frames.shape[0], frames.shape[1]))

    def submit_batch(self, first_utt_id: int, pcm: np.ndarray, offsets: np.ndarray, sample_rate: int) -> int:
        """Whole batch in one float32 buffer; zero-copy when `pcm` is page-locked (keep it alive until sync())."""
        assert pcm.dtype in (np.float32, np.int16) and pcm.flags.c_contiguous
        offsets = np.ascontiguousarray(offsets, np.int64)
        self._keep.append((pcm, offsets))
        if pcm.dtype == np.int16:      # crosses PCIe as int16 (half the bytes), converted on the device
            return self._check(self._lib.fa_submit_pcm_i16_batch(self._h, first_utt_id, pcm.ctypes.data, offsets.ctypes.data,
                                                                 offsets.size - 1, sample_rate))
        return self._check(self._lib.fa_submit_pcm_batch(self._h, first_utt_id, pcm.ctypes.data, offsets.ctypes.data,
                                                         offsets.size - 1, sample_rate))

    def set_truncate(self, on: bool):
        """False: the submitted utterances are prefixes of streams still running -- no segment_truncate at their end."""
        self._check(self._lib.fa_set_truncate(self._h, 1 if on else 0))

    def set_pipeline(self, n_sub: int):
        self._check(self._lib.fa_set_pipeline(self._h, n_sub))

    def set_spectrum_sink(self, dst: np.ndarray | None):
        if dst is None:
            self._sink = None
            self._check(self._lib.fa_set_spectrum_sink(self._h, None, 0))
        else:
            assert dst.dtype == self.spectrum_dtype and dst.flags.c_contiguous and dst.shape[1] == self.cfg.fft_size // 2
            self._sink = dst
            self._check(self._lib.fa_set_spectrum_sink_raw(self._h, dst.ctypes.data, dst.shape[0]))

    def run(self):
        self._check(self._lib.fa_run(self._h))

    def sync(self):
        self._check(self._lib.fa_sync(self._h))
        self._keep.clear()                          # the H2D copies of the batch are done

    def upload(self):
        self._check(self._lib.fa_upload(self._h))

    def run_resident(self):
        self._check(self._lib.fa_run_resident(self._h))

    def download(self):
        self._check(self._lib.fa_download(self._h))

    def stage_times(self):
        ms = (C.c_float * 5)()
        self._check(self._lib.fa_stage_times(self._h, ms))
        return dict(zip(("spectrum", "peaks", "segment", "features", "total"), [float(x) for x in ms]))

    def spectrum_split_times(self):
        """The spectrum stage of the last (serial) run kernel by kernel: FFT magnitudes, smoothing + dB + bands (ms)."""
        ms = (C.c_float * 2)()
        self._check(self._lib.fa_spectrum_split_times(self._h, ms))
        return {"fft": float(ms[0]), "smooth_bands": float(ms[1])}

    @property
    def stream_fixups(self) -> int:
        """Chunks the verification pass of the chunk-parallel smoothing had to recompute in the last run (stream mode)."""
        return self._check(self._lib.fa_stream_fixups(self._h, 0))

    @property
    def control_fixups(self) -> int:
        """Chunks of the control scan that had to be rescanned from the true state in the last run (stream mode)."""
        return self._check(self._lib.fa_stream_fixups(self._h, 1))

    @property
    def k3_redos(self) -> int:
        """Utterances (epochs in stream mode) the fast segment-scan kernel handed back to the general one in the last run
        (more than 64 live tracks or 32 accepted peaks in a frame); > 0 only means extra work, never a different result."""
        return self._check(self._lib.fa_stream_fixups(self._h, 2))

    @property
    def launches(self) -> int:
        return self._lib.fa_launch_count(self._h)

    # -- results --
    def counts(self, utt_id: int | None = None) -> dict:
        c = FaCounts()
        if utt_id is None:
            self._check(self._lib.fa_total_counts(self._h, C.byref(c)))
        else:
            self._check(self._lib.fa_result_counts(self._h, utt_id, C.byref(c)))
        return {k: getattr(c, k) for k, _ in FaCounts._fields_}

    def counts_table(self) -> np.ndarray:
        """fa_counts of every utterance of the batch in submission order, as a structured array (one call)."""
        n = self._lib.fa_num_utterances(self._h)
        out = np.zeros(max(n, 0), COUNTS_DTYPE)
        k = self._check(self._lib.fa_copy_counts_table(self._h, out.ctypes.data, out.size))
        return out[:k]

    def feature_table(self) -> np.ndarray:
        """The dense feature rows of the whole batch ([rows, 53] or [rows, 264]), utterance after utterance."""
        c = self.counts()
        if self.cfg.output_level == 11:
            return self._rows(self._lib.fa_copy_utterance_features, None, c["feature_rows"], (N_UTT_FEATURES,), np.float64)
        if self.cfg.output_level == 12:
            return self._rows(self._lib.fa_copy_curve_features, None, c["feature_rows"], (N_CURVE_FEATURES,), np.float64)
        return self._rows(self._lib.fa_copy_features, None, c["feature_rows"], (N_FEATURES,), np.float64)

    def _rows(self, fn, utt_id, nrows, shape_tail, dtype):
        out = np.empty((nrows,) + shape_tail, dtype)
        n = self._check(fn(self._h, -1 if utt_id is None else utt_id, out.ctypes.data, nrows))
        return out[:n]

    @property
    def spectrum_dtype(self):
        """Element type of the spectrum rows: float32 dB, uint8 (getByteFrequencyData) or float16 (cfg.spectrum_format)."""
        return np.dtype((np.float32, np.uint8, np.float16)[self.cfg.spectrum_format])

    def spectrum(self, utt_id: int | None = None) -> np.ndarray:
        c = self.counts(utt_id)
        return self._rows(self._lib.fa_copy_spectrum_raw, utt_id, c["frames"], (self.cfg.fft_size // 2,), self.spectrum_dtype)

    def frames(self, utt_id: int | None = None) -> np.ndarray:
        c = self.counts(utt_id)
        return self._rows(self._lib.fa_copy_frames, utt_id, c["frames"], (c["bands"],), np.uint32)

    def gsum(self, utt_id: int | None = None) -> np.ndarray:
        c = self.counts(utt_id)
        return self._rows(self._lib.fa_copy_gsum, utt_id, c["frames"], (), np.float64)

    def peak_candidates(self, utt_id: int):
        c = self.counts(utt_id)
        maxp = C.c_int32(0)
        B = c["bands"]
        packed = np.zeros((c["frames"], B // 2 + 4), np.uint32)
        cnt = np.zeros(c["frames"], np.int32)
        self._check(self._lib.fa_copy_peak_candidates(self._h, utt_id, packed.ctypes.data, cnt.ctypes.data, c["frames"],
                                                      C.byref(maxp)))
        assert maxp.value == packed.shape[1]
        return packed, cnt

    def result(self, utt_id: int | None = None) -> UtteranceResult:
        c = self.counts(utt_id)
        L = self._lib
        segs = self._rows(L.fa_copy_segments, utt_id, c["segments"], (), SEG_DTYPE)
        lvl3 = self.cfg.output_level == 3
        fm = self._rows(L.fa_copy_formants, utt_id, 0 if lvl3 else c["formant_rows"], (9,), np.float32)
        en = self._rows(L.fa_copy_energy, utt_id, 0 if lvl3 else c["formant_rows"], (3,), np.float32)
        sy = self._rows(L.fa_copy_syllables, utt_id, c["syllables"], (), SYL_DTYPE)
        if self.cfg.output_level == 3:      # ranked tracks (headers in the syllable table) + their points
            tp = self._rows(L.fa_copy_track_points, utt_id, c["formant_rows"], (), TRACK_POINT_DTYPE)
            return UtteranceResult(c, segs, np.zeros((0, 9), np.float32), np.zeros((0, 3), np.float32), sy, np.zeros((0, N_FEATURES)),
                                   None, tp)
        if self.cfg.output_level == 11:
            ut = self._rows(L.fa_copy_utterance_features, utt_id, c["feature_rows"], (N_UTT_FEATURES,), np.float64)
            return UtteranceResult(c, segs, fm, en, sy, np.zeros((0, N_FEATURES)), ut)
        if self.cfg.output_level == 12:
            ft = self._rows(L.fa_copy_curve_features, utt_id, c["feature_rows"], (N_CURVE_FEATURES,), np.float64)
        else:
            ft = self._rows(L.fa_copy_features, utt_id, c["feature_rows"], (N_FEATURES,), np.float64)
        return UtteranceResult(c, segs, fm, en, sy, ft)
