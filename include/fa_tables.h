/*
 * fa_tables.h -- host-side constant tables of the canonical float32 front end (stage 1 / 1b).
 *
 * The reference's own spectrum stage is an un-vendored AudioWorklet fetched from a CDN
 * (/root/reference/dist/main.js:2@B6480); BASELINE.json replaces it by W3C AnalyserNode semantics.
 * The arithmetic below is therefore builder-defined (DESIGN.md "Front-end spec"); what matters is
 * that the CPU oracle and the CUDA kernels use THE SAME tables, so both include this header.
 * Pure C99, host only.  Tables are computed in double with libm and rounded once to float32.
 */
#ifndef FA_TABLES_H_
#define FA_TABLES_H_

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "fa_b200.h"

#ifndef FA_PI
#define FA_PI 3.14159265358979323846
#endif

/* JS Math.round(sr * step / 1000): halves round up. */
static inline int fa_tab_hop(int sample_rate, double step_ms) {
  double h = floor((double)sample_rate * step_ms / 1000.0 + 0.5);
  return h < 1.0 ? 1 : (int)h;
}

static inline int fa_tab_bands(const fa_config* c) {
  return c->spec_type == FA_SPEC_MEL ? c->n_mel_bins : c->n_fft_bins;
}

static inline int fa_tab_log2(int n) {
  int l = 0;
  while ((1 << l) < n) l++;
  return l;
}

/* Blackman window, alpha = 0.16 (W3C Web Audio API, AnalyserNode). w has N entries. */
static inline void fa_tab_window(int N, float* w) {
  for (int n = 0; n < N; n++) {
    double x = (double)n / (double)N;
    w[n] = (float)(0.42 - 0.5 * cos(2.0 * FA_PI * x) + 0.08 * cos(4.0 * FA_PI * x));
  }
}

/* exp(-2*pi*i*j/P) for j in [0, count), as interleaved (re, im); exact on the axes. */
static inline void fa_tab_unit_roots(int P, int count, float* out) {
  for (int j = 0; j < count; j++) {
    double re, im;
    int q = (int)(((int64_t)4 * j) % P == 0 ? ((int64_t)4 * j) / P : -1);
    if (q >= 0) {
      q &= 3;
      re = q == 0 ? 1.0 : q == 2 ? -1.0 : 0.0;
      im = q == 1 ? -1.0 : q == 3 ? 1.0 : 0.0;
    } else {
      double a = 2.0 * FA_PI * (double)j / (double)P;
      re = cos(a);
      im = -sin(a);
    }
    out[2 * j] = (float)re;
    out[2 * j + 1] = (float)im;
  }
}

/* FFT twiddles of the M-point complex transform: tw[j] = W_M^j, j in [0, M/2). */
static inline void fa_tab_fft_twiddles(int M, float* tw) { fa_tab_unit_roots(M, M / 2, tw); }

/* Real-FFT split twiddles: ws[k] = W_{2M}^k for k in [0, M); entries above M/2 are defined by the
 * mirror rule ws[M-k] = (-re, im) of ws[k] so that a kernel can derive both from one load. */
static inline void fa_tab_split_twiddles(int M, float* ws) {
  fa_tab_unit_roots(2 * M, M / 2 + 1, ws);
  for (int k = M / 2 + 1; k < M; k++) {
    ws[2 * k] = -ws[2 * (M - k)];
    ws[2 * k + 1] = ws[2 * (M - k) + 1];
  }
}

/* Sparse band matrix of the adapter (stage 1b): band m = sum_i w[off[m]+i] * lin[k0[m]+i].
 * Returns the number of weights (caller frees *weights, *k0, *cnt, *off with free()). */
typedef struct fa_bandmat {
  int bands;
  int n_weights;
  int max_taps;
  int* k0;
  int* cnt;
  int* off;
  float* w;
} fa_bandmat;

static inline double fa_tab_hz2mel(double f) { return 2595.0 * log10(1.0 + f / 700.0); }
static inline double fa_tab_mel2hz(double m) { return 700.0 * (pow(10.0, m / 2595.0) - 1.0); }

static inline int fa_tab_bandmat(const fa_config* c, int sample_rate, fa_bandmat* bm) {
  const int N = c->fft_size, half = N / 2, B = fa_tab_bands(c);
  bm->bands = B;
  bm->k0 = (int*)calloc((size_t)B, sizeof(int));
  bm->cnt = (int*)calloc((size_t)B, sizeof(int));
  bm->off = (int*)calloc((size_t)B + 1, sizeof(int));
  bm->w = NULL;
  bm->n_weights = 0;
  bm->max_taps = 0;
  if (!bm->k0 || !bm->cnt || !bm->off) return -1;
  const double df = (double)sample_rate / (double)N;
  /* two passes: count, then fill */
  for (int pass = 0; pass < 2; pass++) {
    int total = 0;
    for (int m = 0; m < B; m++) {
      int first = -1, n = 0;
      if (c->spec_type == FA_SPEC_MEL) {
        const double mlo = fa_tab_hz2mel(c->f_min), mhi = fa_tab_hz2mel(c->f_max);
        const double step = (mhi - mlo) / (double)(B + 1);
        const double left = fa_tab_mel2hz(mlo + step * m), centre = fa_tab_mel2hz(mlo + step * (m + 1)),
                     right = fa_tab_mel2hz(mlo + step * (m + 2));
        int ka = (int)floor(left / df), kb = (int)ceil(right / df);
        if (ka < 1) ka = 1;
        if (kb > half - 1) kb = half - 1;
        for (int k = ka; k <= kb; k++) {
          const double f = k * df;
          double wgt = 0.0;
          if (f > left && f < right) wgt = f <= centre ? (f - left) / (centre - left) : (right - f) / (right - centre);
          if (wgt > 0.0) {
            if (first < 0) first = k;
            /* keep the support contiguous */
            if (pass) bm->w[total + (k - first)] = (float)wgt;
            n = k - first + 1;
          }
        }
      } else {
        const double lo = c->f_max * (double)m / (double)B, hi = c->f_max * (double)(m + 1) / (double)B;
        int ka = (int)ceil(lo / df), kb = (int)ceil(hi / df) - 1; /* lo <= f_k < hi */
        if (ka < 0) ka = 0;
        if (kb > half - 1) kb = half - 1;
        if (kb < ka) { /* empty sub-band: nearest bin to its centre */
          int kc = (int)floor(0.5 * (lo + hi) / df + 0.5);
          if (kc > half - 1) kc = half - 1;
          ka = kb = kc;
        }
        first = ka;
        n = kb - ka + 1;
        if (pass)
          for (int k = ka; k <= kb; k++) bm->w[total + (k - ka)] = (float)(1.0 / (double)n);
      }
      if (first < 0) { first = 0; n = 0; }
      bm->k0[m] = first;
      bm->cnt[m] = n;
      bm->off[m] = total;
      if (n > bm->max_taps) bm->max_taps = n;
      total += n;
    }
    bm->off[B] = total;
    if (!pass) {
      bm->n_weights = total;
      bm->w = (float*)calloc((size_t)(total > 0 ? total : 1), sizeof(float));
      if (!bm->w) return -1;
    }
  }
  return bm->n_weights;
}

static inline void fa_tab_bandmat_free(fa_bandmat* bm) {
  free(bm->k0); free(bm->cnt); free(bm->off); free(bm->w);
  bm->k0 = bm->cnt = bm->off = NULL; bm->w = NULL;
}

/* float32 constants of the adapter */
static inline float fa_tab_gain(const fa_config* c) {
  double ms = c->mag_scale > 0.0 ? c->mag_scale : (double)c->fft_size;
  return (float)(ms * c->pre_norm_gain);
}

static inline int fa_tab_valid_fft(int N) { return N >= 256 && N <= 16384 && (N & (N - 1)) == 0; }

#endif /* FA_TABLES_H_ */
