"""Seeded uint32 frame generators shared by the tests and by the fixture scripts in this directory."""
from __future__ import annotations

import numpy as np


def adversarial_frames(seed: int, B: int, F: int) -> np.ndarray:
    """Smooth random spectra with moving bumps, silences and huge / tiny amplitudes."""
    rng = np.random.default_rng(seed)
    bins = np.arange(B)
    fr = np.zeros((F, B))
    centres = rng.uniform(8, 0.65 * B, size=6)
    for t in range(F):
        centres += rng.normal(0, 0.7, size=6)
        amp = rng.choice([0.0, 1.0, 1.0, 1.0]) * 10 ** rng.uniform(1, 6)
        for c in centres:
            fr[t] += amp * rng.uniform(0.2, 1) * np.exp(-0.5 * ((bins - c) / rng.uniform(0.8, 3)) ** 2)
        fr[t] += rng.uniform(0, 3, size=B)
    return np.rint(fr).astype(np.uint32)


def voiced_random_frames(seed: int, B: int, F: int) -> np.ndarray:
    """Runs of voiced frames (5-8 drifting, jumping, merging bumps with one dominant peak, amplitudes over four decades,
    occasional flat tops and exact ties) separated by pauses of random length and noise level: frames that DO open and
    close segments, exercise the adaptive gate and the track matcher, and sometimes hit the dropped-segment quirk."""
    rng = np.random.default_rng(seed)
    bins = np.arange(B)
    out = np.zeros((F, B))
    t = 0
    while t < F:
        run = int(rng.integers(3, 70))
        nb = int(rng.integers(5, 9))
        centres = np.sort(rng.uniform(8, 0.68 * B, size=nb))
        widths = rng.uniform(0.7, 2.5, size=nb)
        level = 10 ** rng.uniform(1.5, 5.5)
        rel = rng.uniform(0.05, 0.4, size=nb)
        rel[int(rng.integers(0, nb))] = 1.0
        for _ in range(run):
            if t >= F:
                break
            centres += rng.normal(0, 0.6, size=nb)
            if rng.random() < 0.05:
                centres[int(rng.integers(0, nb))] += rng.choice([-12.0, -5.0, 5.0, 12.0])
            level *= 10 ** rng.normal(0, 0.08)
            e = rng.uniform(0, 2.5, size=B) * rng.choice([0.0, 1.0, 4.0])
            for c, w, r in zip(centres, widths, rel):
                e += level * r * rng.uniform(0.7, 1.0) * np.exp(-0.5 * ((bins - c) / w) ** 2)
            if rng.random() < 0.1:
                e = np.floor(e / 16) * 16          # plateaus and ties
            if rng.random() < 0.06:
                e[:] = 0                            # a one-frame drop-out inside the run
            out[t] = e
            t += 1
        pause = int(rng.integers(1, 16))
        floor = rng.choice([0.0, 1.0, 3.0, 30.0])
        for _ in range(pause):
            if t >= F:
                break
            out[t] = rng.uniform(0, 1, size=B) * floor
            t += 1
    return np.rint(np.minimum(out, 4.0e9)).astype(np.uint32)


def dropped_segment_frames(B: int = 128) -> np.ndarray:
    """voiced, pause, voiced, voiced, then a long pause -> len 3 but a track point sits at frame 3: straighten_formants
    throws inside the Promise executor and the segment is dropped after seg_ci.push."""
    bins = np.arange(B)

    def voiced(c, amps):
        e = np.zeros(B)
        for a, cc in zip(amps, c):
            e += a * np.exp(-0.5 * ((bins - cc) / 1.5) ** 2)
        return np.rint(e).astype(np.uint32)

    v = voiced([12, 30, 50, 70, 85], [8000, 40000, 8000, 8000, 8000])
    z = np.zeros(B, np.uint32)
    return np.stack([v, z, v, v] + [z] * 10 + [v] * 12 + [z] * 10)


def dense_peak_frames(seed: int, B: int, F: int, spacing: float) -> np.ndarray:
    """Capacity stress: voiced runs of comb spectra -- a peak every `spacing` bins, all far above the gate, positions
    re-drawn every frame (jitter of +-1 bin, whole comb shifted now and then) so that most peaks found no track and start a
    new one.  spacing 6 keeps ~20 peaks per frame and drives the live tracks (alive for four frames) past the 64 slots of the
    fast K3 kernel; spacing 3 reaches ~40 peaks per frame and more live tracks than any kernel holds (FA_ERR_CAPACITY)."""
    rng = np.random.default_rng(seed)
    bins = np.arange(B)
    out = np.zeros((F, B))
    t = 0
    while t < F:
        run = int(rng.integers(12, 40))
        level = 10 ** rng.uniform(3.0, 5.0)
        phase = rng.uniform(0, spacing)
        dom = rng.uniform(10, 0.6 * B)
        for _ in range(run):
            if t >= F:
                break
            if rng.random() < 0.3:
                phase = rng.uniform(0, spacing)
            centres = np.arange(phase + 2, B - 2, spacing) + rng.integers(-1, 2, size=len(np.arange(phase + 2, B - 2, spacing)))
            e = rng.uniform(0, 2, size=B)
            for c in centres:
                e += level * rng.uniform(0.3, 0.6) * np.exp(-0.5 * ((bins - c) / 0.6) ** 2)
            e += 3 * level * np.exp(-0.5 * ((bins - dom) / 0.8) ** 2)
            out[t] = e
            t += 1
        for _ in range(int(rng.integers(9, 14))):
            if t >= F:
                break
            out[t] = rng.uniform(0, 1, size=B)
            t += 1
    return np.rint(out).astype(np.uint32)
