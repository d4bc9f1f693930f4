"""
TEST INFRASTRUCTURE ONLY (oracle). Not imported by the product package.

Literal, statement-by-statement Python transliteration of the reference's minified hot-path modules
inside /root/reference/dist/main.js (line 2; @B = byte offset from start of file):

  * module 0 "stats"      @B1065-B2714   (readable twin: /root/reference/src/stats.js:29-64)
  * module 3 "segmentor"  @B23403-B31782
  * module 4 "formants"   @B31782-B38281  (make_coeffs / polyfit, level 12, excluded)

Variable names follow the minified identifiers so the text can be diffed by eye against the bundle.
JS Number == Python float (IEEE double); Float32Array == numpy float32 rows; parseInt == trunc;
Math.log10 / Math.pow go through include/fa_jsmath.h (fdlibm, what V8 ships) via a tiny C shim.

It exists to cross-check the faster C restatement (oracle/fa_oracle.c); both are restatements, no JS
engine exists in this image, so parity with the real reference remains "unpinned" (see DESIGN.md).
"""
from __future__ import annotations

import math
from decimal import Decimal, ROUND_HALF_UP

import numpy as np

from .. import jsmath

log10 = jsmath.log10
jspow = jsmath.pow


def parseInt(x: float) -> float:
    return float(math.trunc(x))


def toFixed3(x: float) -> str:
    """Number.prototype.toFixed(3): exact decimal value, ties go up (ECMA-262 21.1.3.3)."""
    return str(Decimal(x).quantize(Decimal("0.001"), rounding=ROUND_HALF_UP))


# ----------------------------------------------------------------------------- module 0 (stats)
def arraySum(arr):
    s = 0.0
    for v in arr:
        s += v
    return s


def arrayMax(arr):
    m = -math.inf
    for v in arr:
        if v > m:
            m = v
    return m


def array_mean_NZ(arr):
    s = 0.0
    nz = 0
    for v in arr:
        if v > 0:
            s += v
            nz += 1
    return s / nz if nz else math.nan  # 0/0 -> NaN in JS


def only_std_NZ(arr):
    mean = array_mean_NZ(arr)
    acc = 0.0
    for v in arr:
        acc += jspow(v - mean, 2)   # (x - t) ** 2 is Math.pow in V8: fdlibm returns x * x exactly, glibc pow may be 1 ulp off
    return math.sqrt(acc / len(arr)) if len(arr) else math.nan


def mean_std_NZ(arr):
    mean = array_mean_NZ(arr)
    acc = 0.0
    for v in arr:
        acc += jspow(v - mean, 2)   # (x - t) ** 2 is Math.pow in V8: fdlibm returns x * x exactly, glibc pow may be 1 ulp off
    return [mean, math.sqrt(acc / len(arr)) if len(arr) else math.nan]


def jsdiv(a, b):
    """JS '/' on doubles (no ZeroDivisionError)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


class JSThrow(Exception):
    """A TypeError the JS would throw (reading a property of undefined)."""


# ----------------------------------------------------------------------------- module 4 (formants)
class Formants:
    i = [3, 4, 6, 9]
    o = 4

    def __init__(self):
        self.l = []  # tracks
        self.s = 0.0
        self.c = 0.0

    # v() @B35919
    def clear_fm(self):
        self.l = []
        self.s = 0.0
        self.c = 0.0

    # _() @B37340
    @staticmethod
    def score(e, t, n, r, a, i, o, l):
        s = 0.0
        if i >= o:
            s = jsdiv(o, i)
        else:
            if not (o > 0):
                return 0
            s = i / o
        if 0 == e:
            return jsdiv(300 * s, t) if s > 0.1 else 0
        if s < 0.001:
            return 0
        if s >= 1:
            s = 10
        elif s < 0.1:
            s = 1
        else:
            s *= 10
        t = 10 - abs(a - r - l)
        if t < 0:
            return 0
        if t < 1:
            t = 1
        i = n
        if i > 10:
            i = 10
        return 10 / e * (t * t + i * s)

    # x() @B35952
    def accumulate_fm(self, e, t, n, r, a):
        l = self.l
        u = len(t)
        if u < 1:
            return
        f = [-1] * u
        d = [0] * u
        self.s += r
        for r_ in range(len(l)):
            a_ = n - l[r_][3]
            if a_ >= 0 and a_ < self.o:
                n_ = len(l[r_][7])
                for o_ in range(u):
                    s_ = abs(l[r_][5] - t[o_][2])
                    if s_ < self.i[a_]:
                        i_ = self.score(a_, s_, n_, l[r_][5], t[o_][2], l[r_][6], e[t[o_][2]], l[r_][4])
                        if i_ > 1 and i_ > d[o_]:
                            d[o_] = i_
                            f[o_] = r_
        for r_ in range(len(l)):
            i_ = [e_ for e_ in range(u) if f[e_] == r_]
            if len(i_) > 0:
                o_ = t[i_[0]][2]
                u_ = e[o_]
                if u_ > a:
                    a_ = t[i_[0]][0]
                    f_ = t[i_[0]][1]
                    for n_ in range(len(i_)):
                        if t[i_[n_]][1] > f_:
                            f_ = t[i_[n_]][1]
                        if t[i_[n_]][0] < a_:
                            a_ = t[i_[n_]][0]
                        if e[t[i_[n_]][2]] > e[o_]:
                            o_ = t[i_[n_]][2]
                    d_ = 0.0
                    for t_ in range(a_, f_ + 1):
                        d_ += e[t_]
                    h = len(l[r_][10])
                    if h >= 3:
                        l[r_][4] = (o_ - l[r_][10][h - 1] + (l[r_][10][h - 2] - l[r_][10][h - 1]) + (l[r_][10][h - 3] - l[r_][10][h - 2])) / 3
                    elif 2 == h:
                        l[r_][4] = (o_ - l[r_][10][h - 1] + (l[r_][10][h - 2] - l[r_][10][h - 1])) / 2
                    elif 1 == h:
                        l[r_][4] = o_ - l[r_][10][h - 1]
                    l[r_][0] = a_
                    l[r_][1] = f_
                    l[r_][2] = n
                    l[r_][3] = n
                    l[r_][5] = o_
                    l[r_][6] = u_
                    l[r_][7].append(n)
                    l[r_][8].append(a_)
                    l[r_][9].append(f_)
                    l[r_][10].append(o_)
                    l[r_][11].append(u_)
                    l[r_][12].append(d_)
                    l[r_][13] += d_
                    l[r_][14] += 1
                    l[r_][15] += d_ * o_
                    l[r_][16] = 0
                    l[r_][17] += f_ - a_ + 1
                    self.s -= d_
                    self.c += d_
        for r_ in range(u):
            if -1 == f[r_]:
                i_ = t[r_][2]
                o_ = e[i_]
                if o_ > a:
                    a_ = t[r_][0]
                    s_ = t[r_][1]
                    c_ = 0.0
                    for t_ in range(a_, s_ + 1):
                        c_ += e[t_]
                    u_ = [a_, s_, n, n, 0, i_, o_, [n], [a_], [s_], [i_], [o_], [c_], c_, 1, c_ * i_, 0, s_ - a_ + 1]
                    l.append(u_)

    # y() @B35670
    def get_ranked_formants(self):
        l = self.l
        e = []
        for t in range(len(l)):
            if l[t][14] >= 2:
                n = jsdiv(l[t][15], l[t][13])
                if n >= 7:
                    r = 0
                    if 0 == len(e):
                        e.append(l[t])
                    else:
                        inserted = False
                        while r < len(e):
                            if jsdiv(e[r][15], e[r][13]) > n:
                                e.insert(r, l[t])
                                inserted = True
                                break
                            r += 1
                        if not inserted and r == len(e):
                            e.append(l[t])
        return e

    # m() @B35074
    @staticmethod
    def straighten_formants(e, t, n):
        r = [np.zeros(9, np.float32) for _ in range(t)]
        a = [np.zeros(3, np.float32) for _ in range(t)]
        i = 0.0
        o = 0
        for t_ in range(len(e)):
            l = jsdiv(e[t_][15], e[t_][13])
            if abs(l - i) > 20 or o < 0:
                i = l
                o += 1
                if o >= 3:
                    break
            for i_ in range(e[t_][14]):
                l_ = o
                s = 3 * l_
                c = 3 * l_ + 1
                u = 3 * l_ + 2
                f = e[t_][10][i_]
                if f > 0:
                    o_ = e[t_][12][i_]
                    d = e[t_][7][i_]
                    h = e[t_][9][i_] - e[t_][8][i_] + 1
                    if d < 0 or d >= t:
                        raise JSThrow("r[d] is undefined")
                    if float(r[d][s]) > n and float(r[d][s]) < f and l_ < 2:
                        if l_ < 3:
                            l_ += 1
                        s = 3 * l_
                        c = 3 * l_ + 1
                        u = 3 * l_ + 2
                    r[d][s] = f
                    r[d][c] = o_
                    r[d][u] = h
                    a[d][0] = np.float32(float(a[d][0]) + f * o_)
                    a[d][1] = np.float32(float(a[d][1]) + o_)
                    a[d][2] = np.float32(float(a[d][2]) + h * o_)
        return [r, a]

    # p() @B34757
    @staticmethod
    def sep_syllables(e, t):
        n = e[0]
        r = e[1]
        a = len(r)
        i = -1
        o = []
        l = []
        s = []
        c = 0
        u = 0
        for e_ in range(a):
            if float(r[e_][1]) > t:
                c = 0
                u += 1
                if i < 0:
                    i = e_
            else:
                c += 1
            if (u > 20 and c > 0) or (u > 10 and c > 1) or (u > 0 and c > 4) or (e_ >= a - 1 and u > 4):
                t_ = e_ - c
                if t_ - i > 1:
                    l.append(n[i:t_])
                    s.append(r[i:t_])
                    o.append([i, t_ - i])
                    i = -1
                    u = 0
        return [o, l, s]

    # u() @B32369
    def formant_features(self, e, t, n):
        a = len(e)
        i = [0] * 3; o = [0] * 3; l = [0] * 3; u = [0] * 3; f = [0] * 3; d = [0] * 3; h = [0] * 3
        p = [0] * 3; m = [0] * 3; g = [0] * 3; y = [0] * 3; v = [0] * 3; x = [0] * 3; _ = [0] * 3; b = [0] * 3
        for n_ in range(3):
            s = False
            c = []; w = []; T = []; k = []; M = []; A = []
            S = 0
            L = 0.0
            for t_ in range(a):
                r = float(e[t_][3 * n_])
                a_ = float(e[t_][3 * n_ + 1])
                if r > 0 and a_ > 0:
                    f_ = float(e[t_][3 * n_ + 2])
                    d_ = 20 * log10(a_)
                    c.append(r * d_); w.append(r); M.append(f_ * d_); T.append(a_); k.append(d_)
                    if s:
                        i_ = r - float(e[t_ - 1][3 * n_])
                        if i_ > 1:
                            l[n_] += i_
                        elif i_ < -1:
                            u[n_] += -1 * i_
                        if a_ > L:
                            L = a_
                            S = 1
                        elif 1 == S and a_ < L / 2:
                            if L > 10:
                                A.append(d_)
                            L = 0.0
                            S = -1
                    if not s:
                        o[n_] += 1
                    s = True
                    i[n_] += 1
                else:
                    s = False
                    S = 0
                    L = 0.0
            if o[n_] > 0:
                e_ = arraySum(T)
                m[n_] = jsdiv(jsdiv(e_, a) * 100, t)
                g[n_] = jsdiv(jsdiv(e_, i[n_]) * 100, t)
                o_ = arraySum(k)
                f[n_] = jsdiv(arraySum(c), o_)
                d[n_] = only_std_NZ(w)
                y[n_] = jsdiv(arraySum(M), o_)
                l_ = mean_std_NZ(k)
                h[n_] = l_[0]
                p[n_] = l_[1]
                v[n_] = len(A)
                if v[n_] > 0:
                    e2 = mean_std_NZ(A)
                    x[n_] = e2[0]
                    b[n_] = e2[1]
                    _[n_] = 100 * (jsdiv(x[n_], jsdiv(o_, len(k))) - 1)
        w_ = []
        w_.append(a); w_.append(math.sqrt(a)); w_.append(jsdiv(self.c, self.s)); w_.append(log10(t)); w_.append(n)
        for e_ in range(3):
            w_ += [f[e_], d[e_], h[e_], p[e_], m[e_], g[e_], y[e_], i[e_], o[e_], l[e_], u[e_], v[e_], x[e_], b[e_], _[e_],
                   jsdiv(100 * i[e_], a)]
        return [float(z) for z in w_]

    # d() @B34407
    def make_syl_features(self, e, t, n):
        return [self.formant_features(seg, t, n) for seg in e[1]]


# ----------------------------------------------------------------------------- module 3 (segmentor)
class Segmentor:
    """reset_segmentation() @B25053 + D/O/C/L/P/I/N.  One instance == one LaunchAudioNodes run."""

    def __init__(self, process_level, spec_bands, plot_len=200, step_ms=15, pause_ms=200, minlen_ms=50,
                 auto_noise_gate=True, voiced_max_dB=150, voiced_min_dB=50, callback=None, test_play=True, labels=()):
        if not spec_bands:
            raise ValueError("Invalid spec_bands")
        self.r = Formants()
        o = self.o = {}
        o["process_level"] = process_level
        o["spec_bands"] = spec_bands
        o["plot_len"] = plot_len
        o["seg_limit_1"] = plot_len - 10
        o["seg_limit_2"] = plot_len - 4
        o["max_voiced_bin"] = parseInt(0.7 * spec_bands)
        o["window_step"] = step_ms / 1e3
        o["seg_breaker"] = pause_ms / step_ms if pause_ms > 2 * step_ms else 250 / step_ms
        o["seg_min_frames"] = parseInt(minlen_ms / step_ms)
        o["current_label"] = list(labels)
        o["play_end"] = False
        o["no_fm_segs"] = 0
        o["c_ci"] = 0
        o["c_started"] = -1
        o["current_frame"] = 0
        o["callbacks_processed"] = 0
        o["auto_noise_gate"] = auto_noise_gate
        o["call_at_end"] = False
        self.l = []  # history (level <= 2)
        self.s = []  # raw ranked tracks
        self.c = []  # formants
        self.f = []  # labels
        self.u = []  # seg_ci
        self.d = []  # seg features
        self.h = []  # syllables
        self.p = []  # syllable features
        if auto_noise_gate:
            self.y = 50.0
            self.v = 2.0
        else:
            self.y = jspow(10, voiced_max_dB / 20)
            self.v = jspow(10, voiced_min_dB / 20)
        self.w = 0
        self.T = 0.0
        self.k = 0
        self.x = self.y
        self._ = self.v
        self.M = test_play
        self.b = callback if not test_play else None
        self.events = []  # (si, label, time, payload) for every callback fired, also when M (recorded separately)
        self.trace = []   # per-frame diagnostics (n, h, p, v_after, c_started_after, c_ci_after)

    # L() @B25649
    def L(self, e=-1):
        o = self.o
        o["c_ci"] = 0
        o["c_started"] = e
        o["no_fm_segs"] = 0
        self.r.clear_fm()

    # C() @B28506
    def C(self, e):
        self.w += 1
        if e > self.y or (self.w > 40 and e > 2 * self.v):
            if e >= self.y:
                self.w = 0
                self.x = self.y = e
            elif e > self.x / 100:
                self.y -= parseInt(self.y / 8)
                self.w = 35
            t = log10(self.y)
            y = self.y
            if t > 7:
                self.v = parseInt(jspow(10, t - 3) / 20)
            elif t > 6:
                self.v = parseInt(jspow(10, t - 3) / 2)
            elif t > 4:
                self.v = parseInt(jspow(10, t - 2) / 2)
            elif t > 2:
                self.v = parseInt(jspow(10, t / 3))
            elif t > 1:
                self.v = parseInt(y / 10)
            else:
                self.v = 1.0
            self._ = self.v
            if self.k > 0 and self.T / self.k < 30 * self.v:
                self.L(0)
                self.k = 0
                self.T = 0.0
            self.T += self.y
            self.k += 1
        elif self.v > 10 and self.v > self._ / 10 and self.w > 20:
            self.v -= parseInt(self._ / 20)
            if self.v < 10:
                self.v = 10.0

    # the body of D() for one frame @B25717
    def _D_frame(self, e):
        o = self.o
        v = self.v
        t = o["c_ci"]
        n = 0; a = 1; i = 0; l = 0; s = 0; c = 0; u = 0
        f = []
        d = 0.0
        h = 2 * v
        p = 0
        g = 0.0
        B = o["spec_bands"]

        def close(upd):
            nonlocal i, s, n, d, h, p
            if upd and e[l] > h:
                h = e[l]
                p = l
            t_ = e[l] / 10
            while i < l and e[i] < t_:
                i += 1
            while s > l and e[s] < t_:
                s -= 1
            f.append([i, s, l])
            n += 1
            d += e[l]

        while a < B:
            g += e[a]
            if e[a] > e[a - 1] and (a < 2 or e[a] > e[a - 2]) and (a < 3 or e[a] > e[a - 3]):
                if -1 == u or 0 == u:
                    if -1 == u and e[l] > v and i <= l and l < s:
                        close(True)
                    i = a - 1
                    l = a
                elif 1 == u:
                    l = a
                u = 1
            elif e[a] < e[a - 1] and (a < 2 or e[a] < e[a - 2]) and (a < 3 or e[a] < e[a - 3]):
                if 1 == u or -1 == u:
                    s = a
                    u = -1
            elif -1 == u:
                c += 1
                if c > 2:
                    c = 0
                    if e[l] > v and i <= l and l < s:
                        close(True)
                    u = 0
            elif 1 == u and e[a] > e[a - 1]:
                l = a
            if a == B - 1 and 1 == u:
                s = a
                l = a
                if e[l] > v and i < l and l <= s:
                    close(False)
            a += 1

        finalize = None
        if o["c_started"] < 0:
            e_ = h * (n - 1) / (d - h) if d > h else 0
            if n > 0 and p > 7 and p < o["max_voiced_bin"] and n > 4 and e_ > 4:
                self.L(0)
                o["c_started"] = 0
            else:
                o["no_fm_segs"] += 1
        if o["c_started"] >= 0:
            if 0 == n or p < 7 or p >= o["max_voiced_bin"] or (n > 3 and jsdiv(d, (g - d)) < 0.1):
                o["no_fm_segs"] += 1
                if o["c_started"] < 2:
                    o["c_started"] -= 1
                elif o["no_fm_segs"] >= o["seg_breaker"]:
                    finalize = self._O(o["c_ci"] + 1)   # Promise executor runs synchronously
                elif o["auto_noise_gate"]:
                    self.C(h)
            else:
                if o["auto_noise_gate"]:
                    self.C(h)
                self.r.accumulate_fm(e, f, t, g, self.v)
                if o["c_started"] < 2:
                    o["c_started"] += 1
                else:
                    o["no_fm_segs"] = 0
        o["c_ci"] += 1
        self.trace.append((n, h, p, self.v, self.y, o["c_started"], o["c_ci"], o["no_fm_segs"]))
        # micro-task of the O() promise: runs before the next frame
        if finalize is not None:
            if finalize == "rejected":
                self.L(-1)
            else:
                self.L(-1)
                if not o["call_at_end"]:
                    self.P()

    # O() @B27088 ; returns 1/0 (resolve value) or "rejected"
    def _O(self, e):
        o = self.o
        r = self.r
        a = e - o["no_fm_segs"]
        lvl = o["process_level"]
        y, v = self.y, self.v
        try:
            if a > o["seg_min_frames"] and o["c_started"] >= 2:
                e_ = o["current_frame"] - a
                i = r.get_ranked_formants()
                if 13 == lvl:
                    self.u.append([e_, a])
                    n = r.straighten_formants(i, a, v)
                    l = r.sep_syllables(n, v)
                    m = r.make_syl_features(l, y, v)
                    self.f.append(list(o["current_label"])); self.s.append(i); self.c.append(n[0]); self.d.append(None)
                    self.h.append(l); self.p.append(m)
                    return 1
                elif 12 == lvl:
                    raise NotImplementedError("level 12 (make_coeffs) is out of scope")
                elif 10 == lvl or 11 == lvl:
                    self.u.append([e_, a])
                    n = r.straighten_formants(i, a, v)
                    l = r.sep_syllables(n, v)
                    self.f.append(list(o["current_label"])); self.s.append(i); self.c.append(n[0]); self.d.append(None)
                    self.h.append(l)
                    return 1
                elif 5 == lvl:
                    self.u.append([e_, a])
                    n = r.straighten_formants(i, a, v)
                    l = r.formant_features(n[0], y, v)
                    self.f.append(list(o["current_label"])); self.s.append(i); self.c.append(n[0]); self.d.append(l)
                    return 1
                elif 4 == lvl:
                    self.u.append([e_, a])
                    n = r.straighten_formants(i, a, v)
                    self.f.append(list(o["current_label"])); self.s.append(i); self.c.append(n[0])
                    return 1
                elif 3 == lvl:
                    self.u.append([e_, a]); self.f.append(list(o["current_label"])); self.s.append(i)
                    return 1
                else:
                    return "rejected"
            return 0
        except JSThrow:
            return "rejected"

    # j() @B31114, V() @B31504
    def _j(self, e):
        o = self.o
        out = []
        for r in range(len(self.h[e][0])):
            out.append([toFixed3((self.u[e][0] + self.h[e][0][r][0]) * o["window_step"]),
                        toFixed3((self.h[e][0][r][1] + 1) * o["window_step"])])
        return out

    def _V(self, e):
        o = self.o
        return [self.u[e][0] * o["window_step"], (self.u[e][1] + 1) * o["window_step"]]

    # P() @B28869 -- records events even when test_play (M) suppresses the user callback
    def P(self):
        o = self.o
        lvl = o["process_level"]

        def fire(*args):
            self.events.append(args)
            if not self.M and self.b:
                self.b(*args)

        if lvl in (12, 13):
            if len(self.p) > len(self.u):
                return True
            while o["callbacks_processed"] < len(self.p):
                o["callbacks_processed"] += 1
                e = o["callbacks_processed"] - 1
                if len(self.p[e]) > 0:
                    fire(e, o["current_label"], self._j(e), self.p[e])
        elif lvl == 10:
            if len(self.h) > len(self.u):
                return True
            while o["callbacks_processed"] < len(self.h):
                o["callbacks_processed"] += 1
                e = o["callbacks_processed"] - 1
                if len(self.h[e][1]) > 0:
                    fire(e, o["current_label"], self._j(e), self.h[e][1])
        elif lvl == 5:
            if len(self.d) > len(self.u):
                return True
            while o["callbacks_processed"] < len(self.d):
                o["callbacks_processed"] += 1
                e = o["callbacks_processed"] - 1
                if len(self.d[e]) > 0:
                    fire(e, o["current_label"], self._V(e), self.d[e])
        elif lvl == 4:
            if len(self.c) > len(self.u):
                return True
            while o["callbacks_processed"] < len(self.c):
                o["callbacks_processed"] += 1
                e = o["callbacks_processed"] - 1
                if len(self.c[e]) > 0:
                    fire(e, o["current_label"], self._V(e), self.c[e])
        elif lvl == 3:
            while o["callbacks_processed"] < len(self.s):
                o["callbacks_processed"] += 1
                e = o["callbacks_processed"] - 1
                if len(self.s[e]) > 0:
                    fire(e, o["current_label"], self.s[e])
        return True

    # I() @B30392
    def spectrum_push(self, e, t=None):
        o = self.o
        if o["spec_bands"] != len(e):
            raise ValueError("Error: bins num mismatch")
        o["current_frame"] += 1
        if o["process_level"] <= 2:
            self.l.append(e)
            if o["process_level"] <= 1:
                del self.l[:-1]
            elif len(self.l) > o["plot_len"]:
                del self.l[0]
            if o["auto_noise_gate"]:
                t_ = arrayMax(e)
                if t_ > self.y:
                    self.y = t_
                    self.w = 0
                    self.x = t_
                elif self.w > o["seg_limit_1"] and self.y > self.x / 4:
                    self.y *= 0.99
                else:
                    self.w += 1
        else:
            self._D_frame([int(z) for z in e])

    # N() @B30800 (the 10 ms timer is immaterial offline)
    def segment_truncate(self):
        o = self.o
        o["play_end"] = True
        res = self._O(o["c_ci"])
        self.L(1)
        if res != "rejected" and o["play_end"]:
            self.P()
