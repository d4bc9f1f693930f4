"""Minimal RIFF/WAVE reader for LaunchAudioNodes(1, <ArrayBuffer>): PCM 8/16/24/32-bit and IEEE float32/64.

The reference hands the encoded file to AudioContext.decodeAudioData (/root/reference/dist/main.js:2@B18693),
which also resamples to the context rate; this build analyses PCM at its native rate (sample-rate
caveat) and down-mixes multi-channel files by averaging, like the mono AnalyserNode input.
"""
from __future__ import annotations

import struct

import numpy as np


class WavError(ValueError):
    pass


def decode_wav(buf: bytes | bytearray | memoryview) -> tuple[np.ndarray, int]:
    b = bytes(buf)
    if len(b) < 12 or b[0:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise WavError("Unable to decode audio data")  # decodeAudioData's EncodingError text
    pos = 12
    fmt = None
    data = None
    while pos + 8 <= len(b):
        cid, size = b[pos:pos + 4], struct.unpack_from("<I", b, pos + 4)[0]
        body = b[pos + 8: pos + 8 + size]
        if cid == b"fmt ":
            if size < 16:
                raise WavError("Unable to decode audio data")
            tag, ch, sr, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == 0xFFFE and size >= 26:  # WAVE_FORMAT_EXTENSIBLE: sub-format GUID starts with the real tag
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise WavError("Unable to decode audio data")
    tag, ch, sr, bits = fmt
    if ch < 1 or sr < 1:
        raise WavError("Unable to decode audio data")
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(data, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(data[: len(data) // 2 * 2], "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            raw = np.frombuffer(data[: len(data) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(data[: len(data) // 4 * 4], "<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise WavError("Unable to decode audio data")
    elif tag == 3:
        if bits == 32:
            x = np.frombuffer(data[: len(data) // 4 * 4], "<f4").astype(np.float32)
        elif bits == 64:
            x = np.frombuffer(data[: len(data) // 8 * 8], "<f8").astype(np.float32)
        else:
            raise WavError("Unable to decode audio data")
    else:
        raise WavError("Unable to decode audio data")
    if ch > 1:
        x = x[: x.size // ch * ch].reshape(-1, ch).mean(axis=1, dtype=np.float32).astype(np.float32)
    return np.ascontiguousarray(x, np.float32), int(sr)


def encode_wav_pcm16(pcm: np.ndarray, sr: int) -> bytes:
    """Test helper: float32 [-1, 1) -> RIFF PCM16 bytes."""
    i16 = np.clip(np.rint(np.asarray(pcm, np.float64) * 32768.0), -32768, 32767).astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(i16)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, sr, sr * 2, 2, 16)
    return hdr + b"data" + struct.pack("<I", len(i16)) + i16
