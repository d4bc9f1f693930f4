/*
 * fa_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the formantanalyzer hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product (webspeechanalyzer_b200/, libfa_b200.so) never does.
 *
 * PARITY STATUS: stages S2 / S2b / S3 / S3b / S3c / S4 / T / CB are PINNED against outputs of the
 * reference's own code: its minified segmentor / formants / stats / utterance modules
 *   - segmentor  /root/reference/dist/main.js:2@B23403-B31782 (inner module 3),
 *   - formants   /root/reference/dist/main.js:2@B31782-B38281 (inner module 4),
 *   - stats      /root/reference/dist/main.js:2@B1065-B2714 == /root/reference/src/stats.js:29-64,
 *   - utterance  /root/reference/dist/main.js:2@B107866-B110125 (inner module 7),
 * were EXECUTED in the build container by oracle/minijs (an ECMAScript-subset interpreter written
 * for this purpose -- the image has no JS engine) on 83 inputs; their outputs are committed as
 * tests/golden/ref_js.json and tests/test_reference_js.py holds this file to them bit for bit.
 * Stage S1 / S1b (spectrum + adapter) remains "parity unpinned" BY NECESSITY: the reference's
 * spectrum stage is an un-vendored CDN worklet (no code, no vectors), so the front end below is a
 * builder-defined AnalyserNode restatement (W3C Web Audio API; DESIGN.md "Front-end spec").
 * Further cross-checks: a literal Python transliteration (oracle/literal/refmodules.py), the
 * invariants the reference pins (row width 53: src/localstore.js:7; feature ranges:
 * dist/nnmodel/1/cats_emotion/model_meta.json), and oracle/run_reference_modules.js for Node.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fPIC -shared -fopenmp (see oracle/build.py).
 * Floating-point contraction MUST be off: every fused multiply-add below is an explicit fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fa_b200.h"
#include "fa_jsmath.h"
#include "fa_curves.h"
#include "fa_tables.h"

#define FAO_API __attribute__((visibility("default")))

/* ============================================================================================
 * Stage 1 / 1b: canonical float32 front end (builder-defined spec)
 * ============================================================================================ */

typedef struct {
  int N, M, logM, B, hop;
  float *win, *tw, *ws;
  fa_bandmat bm;
  float gain, tau, omt, inv2N;
  float* emph; /* [B] (float)(m * high_f_emph) */
} fao_plan;

static int fao_plan_init(fao_plan* p, const fa_config* c, int sr) {
  memset(p, 0, sizeof(*p));
  if (!fa_tab_valid_fft(c->fft_size)) return -1;
  p->N = c->fft_size;
  p->M = p->N / 2;
  p->logM = fa_tab_log2(p->M);
  p->B = fa_tab_bands(c);
  p->hop = fa_tab_hop(sr, c->window_step_ms);
  p->win = (float*)malloc(sizeof(float) * p->N);
  p->tw = (float*)malloc(sizeof(float) * p->M);      /* M/2 complex */
  p->ws = (float*)malloc(sizeof(float) * 2 * p->M);  /* M complex */
  p->emph = (float*)malloc(sizeof(float) * p->B);
  fa_tab_window(p->N, p->win);
  fa_tab_fft_twiddles(p->M, p->tw);
  fa_tab_split_twiddles(p->M, p->ws);
  if (fa_tab_bandmat(c, sr, &p->bm) < 0) return -1;
  p->gain = fa_tab_gain(c);
  p->tau = (float)c->smoothing;
  p->omt = (float)(1.0 - c->smoothing);
  p->inv2N = (float)(1.0 / (2.0 * (double)p->N));
  for (int m = 0; m < p->B; m++) p->emph[m] = (float)((double)m * c->high_f_emph);
  return 0;
}

static void fao_plan_free(fao_plan* p) {
  free(p->win); free(p->tw); free(p->ws); free(p->emph);
  fa_tab_bandmat_free(&p->bm);
}

static unsigned fao_bitrev(unsigned x, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1u); x >>= 1; }
  return r;
}

/* One frame: windowed samples xw[N] -> |X[k]|/N for k in [0, M).  Canonical DAG:
 * packed real FFT, radix-2 decimation-in-time, 6-FMA butterflies, table twiddles. */
static void fao_frame_mag(const fao_plan* p, const float* xw, float* re, float* im, float* mag) {
  const int M = p->M;
  for (int q = 0; q < M; q++) {
    unsigned n = fao_bitrev((unsigned)q, p->logM);
    re[q] = xw[2 * n];
    im[q] = xw[2 * n + 1];
  }
  for (int s = 1; s <= p->logM; s++) {
    const int half = 1 << (s - 1), stride = M >> s;
    for (int base = 0; base < M; base += 2 * half) {
      for (int k = 0; k < half; k++) {
        const float wr = p->tw[2 * (k * stride)], wi = p->tw[2 * (k * stride) + 1];
        const int ia = base + k, ib = ia + half;
        const float ar = re[ia], ai = im[ia], br = re[ib], bi = im[ib];
        const float nr = fmaf(wr, br, fmaf(-wi, bi, ar));
        const float ni = fmaf(wr, bi, fmaf(wi, br, ai));
        re[ia] = nr;
        im[ia] = ni;
        re[ib] = fmaf(2.0f, ar, -nr);
        im[ib] = fmaf(2.0f, ai, -ni);
      }
    }
  }
  for (int k = 0; k < M; k++) {
    const int kb = (M - k) & (M - 1);
    const float ar = re[k], ai = im[k], br = re[kb], bi = im[kb];
    const float sr_ = ar + br, si = ai - bi, dr = ar - br, di = ai + bi;
    const float wr = p->ws[2 * k], wi = p->ws[2 * k + 1];
    const float pp = wi * di, qq = wi * dr;
    const float tr = fmaf(wr, dr, -pp), ti = fmaf(wr, di, qq);
    const float xr = sr_ + ti, xi = si - tr;
    const float m2 = fmaf(xr, xr, xi * xi);
    mag[k] = sqrtf(m2) * p->inv2N;
  }
}

static uint32_t fao_to_u32(float b) {
  if (!(b > 0.0f)) return 0u; /* also NaN */
  if (b >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)rintf(b); /* round-half-even (default rounding mode) */
}

FAO_API int fao_hop(const fa_config* c, int sr) { return fa_tab_hop(sr, c->window_step_ms); }
FAO_API int fao_bands(const fa_config* c) { return fa_tab_bands(c); }
FAO_API int fao_num_frames(const fa_config* c, int sr, int64_t n) { return (int)(n / fa_tab_hop(sr, c->window_step_ms)); }

/* PCM -> per frame: dB spectrum [F][N/2] (nullable), smoothed linear magnitude [F][N/2] (nullable),
 * uint32 band frames [F][B] (nullable).  Returns F, or < 0. */
FAO_API int fao_frontend(const fa_config* c, const float* pcm, int64_t n, int sr, float* spec_db, float* smooth,
                         uint32_t* frames) {
  fao_plan p;
  if (fao_plan_init(&p, c, sr) != 0) return -1;
  const int N = p.N, M = p.M, B = p.B;
  const int F = (int)(n / p.hop);
  float* xw = (float*)malloc(sizeof(float) * N);
  float* re = (float*)malloc(sizeof(float) * M);
  float* im = (float*)malloc(sizeof(float) * M);
  float* mag = (float*)malloc(sizeof(float) * M);
  float* xs = (float*)calloc((size_t)M, sizeof(float));
  float* lin = (float*)malloc(sizeof(float) * M);
  const float lo_db = (float)c->min_db, hi_db = (float)c->max_db;
  for (int t = 0; t < F; t++) {
    const int64_t end = (int64_t)(t + 1) * p.hop, beg = end - N;
    for (int i = 0; i < N; i++) {
      const int64_t j = beg + i;
      const float x = j >= 0 ? pcm[j] : 0.0f;
      xw[i] = x * p.win[i];
    }
    fao_frame_mag(&p, xw, re, im, mag);
    for (int k = 0; k < M; k++) {
      xs[k] = fmaf(p.tau, xs[k], p.omt * mag[k]);
      if (smooth) smooth[(size_t)t * M + k] = xs[k];
      if (spec_db) {
        float d = 20.0f * log10f(xs[k]);
        if (c->clamp_db) d = d < lo_db ? lo_db : (d > hi_db ? hi_db : d);
        spec_db[(size_t)t * M + k] = d;
      }
      lin[k] = xs[k] * p.gain;
      if (c->spec_type == FA_SPEC_POWER) lin[k] = lin[k] * lin[k];
    }
    if (frames) {
      for (int m = 0; m < B; m++) {
        float acc = 0.0f;
        const float* w = p.bm.w + p.bm.off[m];
        const int k0 = p.bm.k0[m];
        for (int i = 0; i < p.bm.cnt[m]; i++) acc = fmaf(w[i], lin[k0 + i], acc);
        if (c->high_f_emph != 0.0) acc = fmaf(acc, p.emph[m], acc);
        frames[(size_t)t * B + m] = fao_to_u32(acc);
      }
    }
  }
  free(xw); free(re); free(im); free(mag); free(xs); free(lin);
  fao_plan_free(&p);
  return F;
}

/* float64 "truth" of the dB spectrum (plain W3C AnalyserNode arithmetic in double, O(N log N)),
 * reported beside the float32 numbers; not part of the parity gate. */
static void fao_fft64(double* re, double* im, int n, int logn) {
  for (int i = 0; i < n; i++) {
    int j = (int)fao_bitrev((unsigned)i, logn);
    if (j > i) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    for (int b = 0; b < n; b += len)
      for (int k = 0; k < len / 2; k++) {
        double a = -2.0 * FA_PI * k / len, wr = cos(a), wi = sin(a);
        int ia = b + k, ib = ia + len / 2;
        double tr = wr * re[ib] - wi * im[ib], ti = wr * im[ib] + wi * re[ib];
        re[ib] = re[ia] - tr; im[ib] = im[ia] - ti;
        re[ia] += tr; im[ia] += ti;
      }
  }
}

FAO_API int fao_frontend_f64(const fa_config* c, const float* pcm, int64_t n, int sr, double* spec_db) {
  if (!fa_tab_valid_fft(c->fft_size)) return -1;
  const int N = c->fft_size, M = N / 2, hop = fa_tab_hop(sr, c->window_step_ms), logN = fa_tab_log2(N);
  const int F = (int)(n / hop);
  double* re = (double*)malloc(sizeof(double) * N);
  double* im = (double*)malloc(sizeof(double) * N);
  double* xs = (double*)calloc((size_t)M, sizeof(double));
  for (int t = 0; t < F; t++) {
    const int64_t end = (int64_t)(t + 1) * hop, beg = end - N;
    for (int i = 0; i < N; i++) {
      const int64_t j = beg + i;
      const double x = j >= 0 ? (double)pcm[j] : 0.0, a = (double)i / N;
      re[i] = x * (0.42 - 0.5 * cos(2 * FA_PI * a) + 0.08 * cos(4 * FA_PI * a));
      im[i] = 0.0;
    }
    fao_fft64(re, im, N, logN);
    for (int k = 0; k < M; k++) {
      const double mg = sqrt(re[k] * re[k] + im[k] * im[k]) / N;
      xs[k] = c->smoothing * xs[k] + (1.0 - c->smoothing) * mg;
      double d = 20.0 * log10(xs[k]);
      if (c->clamp_db) d = d < c->min_db ? c->min_db : (d > c->max_db ? c->max_db : d);
      spec_db[(size_t)t * M + k] = d;
    }
  }
  free(re); free(im); free(xs);
  return F;
}

/* ============================================================================================
 * Stages 2-4: segmentor + formants + features (restatement of the reference)
 * ============================================================================================ */

typedef struct { int *d; int n, cap; } ivec;
typedef struct { double *d; int n, cap; } dvec;
static void ivec_push(ivec* v, int x) {
  if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 8; v->d = (int*)realloc(v->d, sizeof(int) * v->cap); }
  v->d[v->n++] = x;
}
static void dvec_push(dvec* v, double x) {
  if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 8; v->d = (double*)realloc(v->d, sizeof(double) * v->cap); }
  v->d[v->n++] = x;
}

/* formant track: fields of l[r] at @B35952 ([0]lo [1]hi [2]frame [3]lastFrame [4]velocity [5]lastBin
 * [6]lastAmp [7]frames[] [8]los[] [9]his[] [10]bins[] [11]amps[] [12]energies[] [13]sumE [14]count
 * [15]sum(E*bin) [17]sum span) */
typedef struct {
  int lo, hi, frame, last_frame;
  double velocity;
  int last_bin;
  double last_amp;
  ivec frames, los, his, bins;
  dvec energies, amps;
  double sum_e, sum_eb;
  int count, sum_span;
} fao_track;

typedef struct { int lo, hi, pk; } fao_peak;

typedef struct {
  /* config-derived (reset_segmentation @B25053) */
  int level, B, max_voiced_bin, seg_min_frames, auto_gate, plot_len;
  double window_step, seg_breaker;
  /* state */
  int current_frame, no_fm_segs, c_ci, c_started;
  double y, v, x, v0, T;
  int w, k;
  /* formants module state */
  fao_track* tr;
  int n_tr, cap_tr;
  double s_energy, c_energy;
  /* diagnostics for the GPU capacity tests: most tracks that could still be matched after one accumulate_fm call (gap < 4,
   * the live-track slots the CUDA kernels need) and most accepted peaks in one call */
  int max_live, max_peaks;
} fao_state;

struct fao_result {
  int F, B, level;
  /* per-frame trace */
  int *tr_n, *tr_p, *tr_cstart, *tr_cci, *tr_nofm;
  double *tr_h, *tr_v, *tr_y;
  /* seg_ci (array u) */
  ivec seg_start, seg_len, seg_stored; /* stored = index into the stores below or -1 (dropped by the throw) */
  /* stores (arrays c/d/h/p share the index) */
  ivec st_len, st_row_off, st_nsyl, st_first_syl;
  dvec st_y, st_v, st_cs;
  float* formants; /* rows of 9 */
  float* energy;   /* rows of 3 */
  int n_rows, cap_rows;
  ivec syl_seg, syl_start, syl_len;
  ivec syl_flag; /* level 12: non-zero where the reference's make_coeffs would have thrown (its row and the later ones of the segment are dropped) */
  dvec features; /* rows of 53: level 5 -> one per stored segment, level 13 -> one per syllable */
  int n_feature_rows;
  /* level 3: the ranked tracks of every stored segment (array s) and their points, in rank / point order */
  ivec trk_seg, trk_first, trk_n;
  ivec pt_frame, pt_lo, pt_hi, pt_bin;
  dvec pt_amp, pt_e;
  /* callback order (P @B28869): indices into seg_ci, in firing order */
  ivec cb_si;
  /* level 11: one 264-dim row per callback = get_utterance_features(u, h) over the stores that exist at that time */
  dvec utt_rows;
  int n_utt_rows;
  int max_live, max_peaks; /* see fao_state */
};
typedef struct fao_result fao_result;

static void track_free(fao_track* t) { free(t->frames.d); free(t->los.d); free(t->his.d); free(t->bins.d); free(t->energies.d); free(t->amps.d); }

/* clear_fm @B35919 */
static void clear_fm(fao_state* st) {
  for (int i = 0; i < st->n_tr; i++) track_free(&st->tr[i]);
  st->n_tr = 0;
  st->s_energy = 0.0;
  st->c_energy = 0.0;
}

/* L() @B25649 */
static void seg_reset(fao_state* st, int started) {
  st->c_ci = 0;
  st->c_started = started;
  st->no_fm_segs = 0;
  clear_fm(st);
}

/* _() @B37340 */
static double fm_score(int gap, double dist, int count, int bin_old, int bin_new, double amp_old, double amp_new,
                       double velocity) {
  double s;
  if (amp_old >= amp_new) s = amp_new / amp_old;
  else {
    if (!(amp_new > 0)) return 0;
    s = amp_old / amp_new;
  }
  if (gap == 0) return s > 0.1 ? 300.0 * s / dist : 0;
  if (s < 0.001) return 0;
  if (s >= 1) s = 10; else if (s < 0.1) s = 1; else s *= 10;
  double t = 10.0 - fabs((double)bin_new - (double)bin_old - velocity);
  if (t < 0) return 0;
  if (t < 1) t = 1;
  int i = count > 10 ? 10 : count;
  return 10.0 / (double)gap * (t * t + (double)i * s);
}

static fao_track* track_new(fao_state* st) {
  if (st->n_tr == st->cap_tr) {
    st->cap_tr = st->cap_tr ? 2 * st->cap_tr : 64;
    st->tr = (fao_track*)realloc(st->tr, sizeof(fao_track) * st->cap_tr);
  }
  fao_track* t = &st->tr[st->n_tr++];
  memset(t, 0, sizeof(*t));
  return t;
}

/* accumulate_fm @B35952 */
static void accumulate_fm(fao_state* st, const uint32_t* e, const fao_peak* pk, int u, int n, double g, double vmin) {
  static const int DIST[4] = {3, 4, 6, 9};
  if (u < 1) return;
  int* owner = (int*)malloc(sizeof(int) * u);
  double* best = (double*)malloc(sizeof(double) * u);
  for (int o = 0; o < u; o++) { owner[o] = -1; best[o] = 0; }
  st->s_energy += g;
  const int ntr = st->n_tr;
  for (int r = 0; r < ntr; r++) {
    const fao_track* t = &st->tr[r];
    int gap = n - t->last_frame;
    if (gap >= 0 && gap < 4) {
      for (int o = 0; o < u; o++) {
        int dist = abs(t->last_bin - pk[o].pk);
        if (dist < DIST[gap]) {
          double sc = fm_score(gap, (double)dist, t->frames.n, t->last_bin, pk[o].pk, t->last_amp, (double)e[pk[o].pk],
                               t->velocity);
          if (sc > 1 && sc > best[o]) { best[o] = sc; owner[o] = r; }
        }
      }
    }
  }
  for (int r = 0; r < ntr; r++) {
    fao_track* t = &st->tr[r];
    int first = -1;
    for (int o = 0; o < u; o++) if (owner[o] == r) { first = o; break; }
    if (first < 0) continue;
    int o_bin = pk[first].pk;
    double amp = (double)e[o_bin];
    if (amp > vmin) {
      int lo = pk[first].lo, hi = pk[first].hi;
      for (int o = first; o < u; o++) if (owner[o] == r) {
        if (pk[o].hi > hi) hi = pk[o].hi;
        if (pk[o].lo < lo) lo = pk[o].lo;
        if (e[pk[o].pk] > e[o_bin]) o_bin = pk[o].pk;
      }
      double E = 0;
      for (int b = lo; b <= hi; b++) E += (double)e[b];
      int h = t->bins.n;
      const int* bn = t->bins.d;
      if (h >= 3) t->velocity = ((double)(o_bin - bn[h - 1] + (bn[h - 2] - bn[h - 1]) + (bn[h - 3] - bn[h - 2]))) / 3;
      else if (h == 2) t->velocity = ((double)(o_bin - bn[h - 1] + (bn[h - 2] - bn[h - 1]))) / 2;
      else if (h == 1) t->velocity = (double)(o_bin - bn[h - 1]);
      t->lo = lo; t->hi = hi; t->frame = n; t->last_frame = n; t->last_bin = o_bin; t->last_amp = amp;
      ivec_push(&t->frames, n); ivec_push(&t->los, lo); ivec_push(&t->his, hi); ivec_push(&t->bins, o_bin);
      dvec_push(&t->energies, E); dvec_push(&t->amps, amp);
      t->sum_e += E; t->count += 1; t->sum_eb += E * (double)o_bin; t->sum_span += hi - lo + 1;
      st->s_energy -= E;
      st->c_energy += E;
    }
  }
  for (int o = 0; o < u; o++) if (owner[o] == -1) {
    int i = pk[o].pk;
    double amp = (double)e[i];
    if (amp > vmin) {
      int lo = pk[o].lo, hi = pk[o].hi;
      double E = 0;
      for (int b = lo; b <= hi; b++) E += (double)e[b];
      fao_track* t = track_new(st);
      t->lo = lo; t->hi = hi; t->frame = n; t->last_frame = n; t->velocity = 0; t->last_bin = i; t->last_amp = amp;
      ivec_push(&t->frames, n); ivec_push(&t->los, lo); ivec_push(&t->his, hi); ivec_push(&t->bins, i);
      dvec_push(&t->energies, E); dvec_push(&t->amps, amp);
      t->sum_e = E; t->count = 1; t->sum_eb = E * (double)i; t->sum_span = hi - lo + 1;
    }
  }
  free(owner); free(best);
  {
    int live = 0;
    for (int r = 0; r < st->n_tr; r++) if (n - st->tr[r].last_frame < 4) live++;
    if (live > st->max_live) st->max_live = live;
    if (u > st->max_peaks) st->max_peaks = u;
  }
}

/* get_ranked_formants @B35670: indices of tracks, ascending mean, stable */
static int ranked_formants(const fao_state* st, int* out) {
  int n = 0;
  for (int t = 0; t < st->n_tr; t++) {
    const fao_track* tk = &st->tr[t];
    if (tk->count >= 2) {
      double m = tk->sum_eb / tk->sum_e;
      if (m >= 7) {
        int r = 0;
        while (r < n) {
          const fao_track* o = &st->tr[out[r]];
          if (o->sum_eb / o->sum_e > m) break;
          r++;
        }
        memmove(out + r + 1, out + r, sizeof(int) * (n - r));
        out[r] = t;
        n++;
      }
    }
  }
  return n;
}

/* straighten_formants @B35074.  Returns 0, or -1 where the JS would throw (frame index >= len). */
static int straighten(const fao_state* st, const int* ranked, int nr, int len, double vmin, float* F, float* Eg) {
  memset(F, 0, sizeof(float) * 9 * (size_t)len);
  memset(Eg, 0, sizeof(float) * 3 * (size_t)len);
  double anchor = 0;
  int slot = 0;
  for (int t = 0; t < nr; t++) {
    const fao_track* tk = &st->tr[ranked[t]];
    double m = tk->sum_eb / tk->sum_e;
    if (fabs(m - anchor) > 20 || slot < 0) {
      anchor = m;
      slot++;
      if (slot >= 3) break;
    }
    for (int i = 0; i < tk->count; i++) {
      int sl = slot;
      int bin = tk->bins.d[i];
      if (bin > 0) {
        double E = tk->energies.d[i];
        int fr = tk->frames.d[i];
        int span = tk->his.d[i] - tk->los.d[i] + 1;
        if (fr < 0 || fr >= len) return -1; /* r[d] undefined -> TypeError */
        float* row = F + 9 * (size_t)fr;
        float* eg = Eg + 3 * (size_t)fr;
        if ((double)row[3 * sl] > vmin && (double)row[3 * sl] < (double)bin && sl < 2) sl++;
        row[3 * sl] = (float)bin;
        row[3 * sl + 1] = (float)E;
        row[3 * sl + 2] = (float)span;
        eg[0] = (float)((double)eg[0] + (double)bin * E);
        eg[1] = (float)((double)eg[1] + E);
        eg[2] = (float)((double)eg[2] + (double)span * E);
      }
    }
  }
  return 0;
}

/* stats (src/stats.js:29-64) */
static double array_mean_nz(const double* a, int n) {
  double s = 0; int c = 0;
  for (int i = 0; i < n; i++) if (a[i] > 0) { s += a[i]; c++; }
  return s / (double)c;
}
static double std_nz(const double* a, int n, double* mean_out) {
  double mean = array_mean_nz(a, n), acc = 0;
  for (int i = 0; i < n; i++) { double d = a[i] - mean; acc += d * d; }
  if (mean_out) *mean_out = mean;
  return sqrt(acc / (double)n);
}
static double arr_sum(const double* a, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i]; return s; }

/* formant_features @B32369 (vector assembly @B33436). F: len rows of 9 float32. */
static void formant_features(const float* F, int len, double ymax, double vmin, double cs, double* out) {
  double cnt[3] = {0}, runs[3] = {0}, up[3] = {0}, down[3] = {0}, fmean[3] = {0}, fstd[3] = {0}, dbm[3] = {0},
         dbs[3] = {0}, e_len[3] = {0}, e_cnt[3] = {0}, span[3] = {0}, nacc[3] = {0}, accm[3] = {0}, accs[3] = {0},
         prom[3] = {0};
  double* c = (double*)malloc(sizeof(double) * 6 * (size_t)(len > 0 ? len : 1));
  double *w = c + len, *T = w + len, *k = T + len, *M = k + len, *A = M + len;
  for (int n = 0; n < 3; n++) {
    int prev = 0, S = 0, m = 0, na = 0;
    double L = 0;
    for (int t = 0; t < len; t++) {
      double r = (double)F[9 * (size_t)t + 3 * n], a = (double)F[9 * (size_t)t + 3 * n + 1];
      if (r > 0 && a > 0) {
        double f = (double)F[9 * (size_t)t + 3 * n + 2], d = 20.0 * fa_js_log10(a);
        c[m] = r * d; w[m] = r; M[m] = f * d; T[m] = a; k[m] = d; m++;
        if (prev) {
          double j = r - (double)F[9 * (size_t)(t - 1) + 3 * n];
          if (j > 1) up[n] += j; else if (j < -1) down[n] += -1 * j;
          if (a > L) { L = a; S = 1; }
          else if (S == 1 && a < L / 2) { if (L > 10) A[na++] = d; L = 0; S = -1; }
        }
        if (!prev) runs[n] += 1;
        prev = 1;
        cnt[n] += 1;
      } else { prev = 0; S = 0; L = 0; }
    }
    if (runs[n] > 0) {
      double e = arr_sum(T, m);
      e_len[n] = e / (double)len * 100 / ymax;
      e_cnt[n] = e / cnt[n] * 100 / ymax;
      double o = arr_sum(k, m);
      fmean[n] = arr_sum(c, m) / o;
      fstd[n] = std_nz(w, m, NULL);
      span[n] = arr_sum(M, m) / o;
      dbs[n] = std_nz(k, m, &dbm[n]);
      nacc[n] = na;
      if (na > 0) {
        accs[n] = std_nz(A, na, &accm[n]);
        prom[n] = 100 * (accm[n] / (o / (double)m) - 1);
      }
    }
  }
  int q = 0;
  out[q++] = len; out[q++] = sqrt((double)len); out[q++] = cs; out[q++] = fa_js_log10(ymax); out[q++] = vmin;
  for (int e = 0; e < 3; e++) {
    out[q++] = fmean[e]; out[q++] = fstd[e]; out[q++] = dbm[e]; out[q++] = dbs[e]; out[q++] = e_len[e];
    out[q++] = e_cnt[e]; out[q++] = span[e]; out[q++] = cnt[e]; out[q++] = runs[e]; out[q++] = up[e];
    out[q++] = down[e]; out[q++] = nacc[e]; out[q++] = accm[e]; out[q++] = accs[e]; out[q++] = prom[e];
    out[q++] = 100 * cnt[e] / (double)len;
  }
  free(c);
}

/* C() @B28506 */
static void noise_gate(fao_state* st, double e) {
  st->w++;
  if (e > st->y || (st->w > 40 && e > 2 * st->v)) {
    if (e >= st->y) { st->w = 0; st->x = st->y = e; }
    else if (e > st->x / 100) { st->y -= fa_js_parse_int(st->y / 8); st->w = 35; }
    double t = fa_js_log10(st->y);
    if (t > 7) st->v = fa_js_parse_int(fa_js_pow(10, t - 3) / 20);
    else if (t > 6) st->v = fa_js_parse_int(fa_js_pow(10, t - 3) / 2);
    else if (t > 4) st->v = fa_js_parse_int(fa_js_pow(10, t - 2) / 2);
    else if (t > 2) st->v = fa_js_parse_int(fa_js_pow(10, t / 3));
    else if (t > 1) st->v = fa_js_parse_int(st->y / 10);
    else st->v = 1;
    st->v0 = st->v;
    if (st->k > 0 && st->T / (double)st->k < 30 * st->v) { seg_reset(st, 0); st->k = 0; st->T = 0; }
    st->T += st->y;
    st->k += 1;
  } else if (st->v > 10 && st->v > st->v0 / 10 && st->w > 20) {
    st->v -= fa_js_parse_int(st->v0 / 20);
    if (st->v < 10) st->v = 10;
  }
}

static void rows_reserve(fao_result* R, int extra) {
  if (R->n_rows + extra > R->cap_rows) {
    R->cap_rows = 2 * (R->n_rows + extra) + 64;
    R->formants = (float*)realloc(R->formants, sizeof(float) * 9 * (size_t)R->cap_rows);
    R->energy = (float*)realloc(R->energy, sizeof(float) * 3 * (size_t)R->cap_rows);
  }
}

/* sep_syllables @B34757: emits [start,len] pairs */
static int sep_syllables(const float* Eg, int len, double vmin, int* starts, int* lens) {
  int start = -1, quiet = 0, loud = 0, n = 0;
  for (int e = 0; e < len; e++) {
    if ((double)Eg[3 * (size_t)e + 1] > vmin) { quiet = 0; loud++; if (start < 0) start = e; }
    else quiet++;
    if ((loud > 20 && quiet > 0) || (loud > 10 && quiet > 1) || (loud > 0 && quiet > 4) || (e >= len - 1 && loud > 4)) {
      int end = e - quiet;
      if (end - start > 1) { starts[n] = start; lens[n] = end - start; n++; start = -1; loud = 0; }
    }
  }
  return n;
}

/* O() @B27088. Returns 1 stored, 0 ignored, -1 rejected (JS throw). */
static int finalize_segment(fao_state* st, fao_result* R, int n_arg) {
  int len = n_arg - st->no_fm_segs;
  if (!(len > st->seg_min_frames && st->c_started >= 2)) return 0;
  if (!(st->level == 3 || st->level == 4 || st->level == 5 || st->level == 10 || st->level == 11 || st->level == 12 ||
        st->level == 13))
    return -1;
  int start = st->current_frame - len;
  int* ranked = (int*)malloc(sizeof(int) * (size_t)(st->n_tr > 0 ? st->n_tr : 1));
  int nr = ranked_formants(st, ranked);
  if (getenv("FAO_DEBUG_SEG")) {
    int np = 0;
    for (int t = 0; t < st->n_tr; t++) np += st->tr[t].count;
    fprintf(stderr, "FAO_SEG T=%d NP=%d len=%d cci=%d nr=%d\n", st->n_tr, np, len, st->c_ci, nr);
    for (int t = 0; t < st->n_tr; t++) {   /* frame labels of a track: increasing, except a stale first label */
      const fao_track* tk = &st->tr[t];
      for (int i = 1; i < tk->frames.n; i++)
        if (tk->frames.d[i] <= tk->frames.d[i - 1])
          fprintf(stderr, "FAO_SEG   track %d point %d: label %d after %d\n", t, i, tk->frames.d[i], tk->frames.d[i - 1]);
    }
  }
  ivec_push(&R->seg_start, start);
  ivec_push(&R->seg_len, len);
  ivec_push(&R->seg_stored, -1);
  int si = R->seg_start.n - 1;
  int rc = 1;
  if (st->level >= 4) {
    rows_reserve(R, len);
    float* F = R->formants + 9 * (size_t)R->n_rows;
    float* Eg = R->energy + 3 * (size_t)R->n_rows;
    if (straighten(st, ranked, nr, len, st->v, F, Eg) != 0) rc = -1;
    else {
      int stored = R->st_len.n;
      R->seg_stored.d[si] = stored;
      ivec_push(&R->st_len, len);
      ivec_push(&R->st_row_off, R->n_rows);
      dvec_push(&R->st_y, st->y);
      dvec_push(&R->st_v, st->v);
      double cs = st->c_energy / st->s_energy;
      dvec_push(&R->st_cs, cs);
      int nsyl = 0;
      ivec_push(&R->st_first_syl, R->syl_seg.n);
      if (st->level == 10 || st->level == 11 || st->level == 12 || st->level == 13) {
        int* ss = (int*)malloc(sizeof(int) * 2 * (size_t)len);
        nsyl = sep_syllables(Eg, len, st->v, ss, ss + len);
        int thrown = 0;   /* level 12: make_coeffs' try / catch returns the rows made before a throw inside numeric */
        for (int i = 0; i < nsyl; i++) {
          ivec_push(&R->syl_seg, stored); ivec_push(&R->syl_start, ss[i]); ivec_push(&R->syl_len, ss[len + i]);
          if (st->level == 12) {
            /* make_coeffs @B34527: one row of 23 per syllable; a syllable the reference would have thrown on (and every later
             * one of the segment) keeps a NaN row and a non-zero flag */
            double row[FA_N_CURVE_FEATURES];
            const int sl = ss[len + i];
            double* work = (double*)malloc(sizeof(double) * 11 * (size_t)(sl > 0 ? sl : 1));
            for (int w = 0; w < 4 && !thrown; w++)
              thrown = fa_curve_fit_one(F + 9 * (size_t)ss[i], Eg + 3 * (size_t)ss[i], sl, w, work, row + fa_curve_slice_offset(w));
            free(work);
            if (thrown) for (int q = 0; q < FA_N_CURVE_FEATURES; q++) row[q] = 0.0 / 0.0;
            for (int q = 0; q < FA_N_CURVE_FEATURES; q++) dvec_push(&R->features, row[q]);
            ivec_push(&R->syl_flag, thrown);
            R->n_feature_rows++;
          }
          if (st->level == 13) {
            double row[FA_N_FEATURES];
            formant_features(F + 9 * (size_t)ss[i], ss[len + i], st->y, st->v, cs, row);
            for (int q = 0; q < FA_N_FEATURES; q++) dvec_push(&R->features, row[q]);
            R->n_feature_rows++;
          }
        }
        free(ss);
      }
      ivec_push(&R->st_nsyl, nsyl);
      if (st->level == 5) {
        double row[FA_N_FEATURES];
        formant_features(F, len, st->y, st->v, cs, row);
        for (int q = 0; q < FA_N_FEATURES; q++) dvec_push(&R->features, row[q]);
        R->n_feature_rows++;
      }
      R->n_rows += len;
    }
  } else {
    /* level 3: raw ranked tracks only; record the segment with an empty store entry */
    int stored = R->st_len.n;
    R->seg_stored.d[si] = stored;
    ivec_push(&R->st_len, 0); ivec_push(&R->st_row_off, R->pt_frame.n); dvec_push(&R->st_y, st->y); dvec_push(&R->st_v, st->v);
    dvec_push(&R->st_cs, st->c_energy / st->s_energy); ivec_push(&R->st_first_syl, R->trk_seg.n); ivec_push(&R->st_nsyl, nr);
    /* s.push(get_ranked_formants()): the track arrays themselves (18 fields, @B35952), kept here as headers + points */
    for (int r = 0; r < nr; r++) {
      const fao_track* t = &st->tr[ranked[r]];
      ivec_push(&R->trk_seg, stored); ivec_push(&R->trk_first, R->pt_frame.n); ivec_push(&R->trk_n, t->count);
      for (int i = 0; i < t->count; i++) {
        ivec_push(&R->pt_frame, t->frames.d[i]); ivec_push(&R->pt_lo, t->los.d[i]); ivec_push(&R->pt_hi, t->his.d[i]);
        ivec_push(&R->pt_bin, t->bins.d[i]); dvec_push(&R->pt_amp, t->amps.d[i]); dvec_push(&R->pt_e, t->energies.d[i]);
      }
    }
  }
  free(ranked);
  return rc;
}

/* v-independent candidate scan (stage 2 on the GPU): emits every peak the reference's scan WOULD
 * close if e[pk] > v held, with the trimmed bounds.  packed = lo | hi<<8 | pk<<16 | lastbin<<24. */
FAO_API int fao_peak_candidates(const uint32_t* e, int B, uint32_t* packed, double* gsum) {
  int n = 0, lo = 0, pk = 0, hi = 0, flat = 0, dir = 0;
  double g = 0;
#define FAO_EMIT(last)                                                    \
  do {                                                                    \
    int l2 = lo, h2 = hi;                                                 \
    double thr = (double)e[pk] / 10;                                      \
    while (l2 < pk && (double)e[l2] < thr) l2++;                          \
    while (h2 > pk && (double)e[h2] < thr) h2--;                          \
    packed[n++] = (uint32_t)l2 | ((uint32_t)h2 << 8) | ((uint32_t)pk << 16) | ((uint32_t)(last) << 24); \
  } while (0)
  for (int a = 1; a < B; a++) {
    g += (double)e[a];
    if (e[a] > e[a - 1] && (a < 2 || e[a] > e[a - 2]) && (a < 3 || e[a] > e[a - 3])) {
      if (dir == -1 || dir == 0) {
        if (dir == -1 && lo <= pk && pk < hi) FAO_EMIT(0);
        lo = a - 1; pk = a;
      } else if (dir == 1) pk = a;
      dir = 1;
    } else if (e[a] < e[a - 1] && (a < 2 || e[a] < e[a - 2]) && (a < 3 || e[a] < e[a - 3])) {
      if (dir == 1 || dir == -1) { hi = a; dir = -1; }
    } else if (dir == -1) {
      flat++;
      if (flat > 2) {
        flat = 0;
        if (lo <= pk && pk < hi) FAO_EMIT(0);
        dir = 0;
      }
    } else if (dir == 1 && e[a] > e[a - 1]) pk = a;
    if (a == B - 1 && dir == 1) {
      hi = a; pk = a;
      if (lo < pk && pk <= hi) FAO_EMIT(1);
    }
  }
#undef FAO_EMIT
  if (gsum) *gsum = g;
  return n;
}

/* one frame of D() @B25717 (literal v-dependent scan). Returns finalisation result or -2 if none. */
static int process_frame(fao_state* st, fao_result* R, const uint32_t* e, int frame_idx) {
  const int B = st->B;
  const double v = st->v;
  int t = st->c_ci, n = 0, i = 0, l = 0, s = 0, c = 0, u = 0, p = 0;
  double d = 0, h = 2 * v, g = 0;
  fao_peak* f = (fao_peak*)malloc(sizeof(fao_peak) * (size_t)B);
#define FAO_CLOSE(upd)                                                  \
  do {                                                                  \
    if ((upd) && (double)e[l] > h) { h = (double)e[l]; p = l; }         \
    double thr = (double)e[l] / 10;                                     \
    while (i < l && (double)e[i] < thr) i++;                            \
    while (s > l && (double)e[s] < thr) s--;                            \
    f[n].lo = i; f[n].hi = s; f[n].pk = l; n++; d += (double)e[l];      \
  } while (0)
  for (int a = 1; a < B; a++) {
    g += (double)e[a];
    if (e[a] > e[a - 1] && (a < 2 || e[a] > e[a - 2]) && (a < 3 || e[a] > e[a - 3])) {
      if (u == -1 || u == 0) {
        if (u == -1 && (double)e[l] > v && i <= l && l < s) FAO_CLOSE(1);
        i = a - 1; l = a;
      } else if (u == 1) l = a;
      u = 1;
    } else if (e[a] < e[a - 1] && (a < 2 || e[a] < e[a - 2]) && (a < 3 || e[a] < e[a - 3])) {
      if (u == 1 || u == -1) { s = a; u = -1; }
    } else if (u == -1) {
      c++;
      if (c > 2) {
        c = 0;
        if ((double)e[l] > v && i <= l && l < s) FAO_CLOSE(1);
        u = 0;
      }
    } else if (u == 1 && e[a] > e[a - 1]) l = a;
    if (a == B - 1 && u == 1) {
      s = a; l = a;
      if ((double)e[l] > v && i < l && l <= s) FAO_CLOSE(0);
    }
  }
#undef FAO_CLOSE
  int fin = -2;
  if (st->c_started < 0) {
    double ratio = d > h ? h * (double)(n - 1) / (d - h) : 0;
    if (n > 0 && p > 7 && p < st->max_voiced_bin && n > 4 && ratio > 4) { seg_reset(st, 0); st->c_started = 0; }
    else st->no_fm_segs++;
  }
  if (st->c_started >= 0) {
    if (n == 0 || p < 7 || p >= st->max_voiced_bin || (n > 3 && d / (g - d) < 0.1)) {
      st->no_fm_segs++;
      if (st->c_started < 2) st->c_started--;
      else if ((double)st->no_fm_segs >= st->seg_breaker) fin = finalize_segment(st, R, st->c_ci + 1);
      else if (st->auto_gate) noise_gate(st, h);
    } else {
      if (st->auto_gate) noise_gate(st, h);
      accumulate_fm(st, e, f, n, t, g, st->v);
      if (st->c_started < 2) st->c_started++; else st->no_fm_segs = 0;
    }
  }
  st->c_ci++;
  if (R->tr_n) {
    R->tr_n[frame_idx] = n; R->tr_p[frame_idx] = p; R->tr_h[frame_idx] = h; R->tr_v[frame_idx] = st->v;
    R->tr_y[frame_idx] = st->y; R->tr_cstart[frame_idx] = st->c_started; R->tr_cci[frame_idx] = st->c_ci;
    R->tr_nofm[frame_idx] = st->no_fm_segs;
  }
  free(f);
  return fin;
}

/* ---- level 11: get_utterance_features @B107983 (inner module 7 of the bundle) ------------------------------------
 * 15 histograms (module-level arrays i,o,l,s,c,u,f,d,h,p,m,g,y,v,x -- sizes below, 264 bins in all) over the segments
 * and syllables stored so far.  `arr[k]++` with k = NaN or k < 0 does not touch a bin: it creates the named property
 * "NaN" / "-1" holding NaN, which w() @B109452 then meets in its for-in sum -- the total becomes NaN, `t > 0` fails and
 * the histogram is returned UN-normalised.  fao_hist.poison models that property. */
typedef struct { double c[40]; int n, poison; } fao_hist;
enum { H_I, H_O, H_L, H_S, H_C, H_U, H_F, H_D, H_H, H_P, H_M, H_G, H_Y, H_V, H_X, H_COUNT };
static const int kHistSize[H_COUNT] = {10, 10, 10, 10, 20, 40, 40, 24, 24, 8, 8, 10, 10, 20, 20};

static void hist_inc(fao_hist* h, double k) {
  if (k != k || k <= -1.0) { h->poison = 1; return; }   /* named property; (-1, 0) truncates to -0 -> index 0 */
  int i = (int)k;
  if (i >= h->n) { h->poison = 1; return; }              /* cannot happen: every index is clamped from above */
  h->c[i] += 1;
}

/* _() @B109020: one segment */
static void utt_segment(fao_hist* H, double seg_len, double n_syl, double gap, double voiced) {
  double a = fa_js_parse_int(10 * seg_len / 150); if (a >= 10) a = 9; hist_inc(&H[H_I], a);
  double c = n_syl; if (c >= 10) c = 9; hist_inc(&H[H_O], c);
  double u = fa_js_parse_int(10 * gap / 150); if (u >= 10) u = 9; hist_inc(&H[H_L], u);
  double f = fa_js_parse_int(2 * (voiced - .3) * 10); if (f >= 10) f = 9; if (f < 0) f = 0; hist_inc(&H[H_S], f);
}

/* b() @B109255: one syllable (e len, t/n/a mean f0 bin/energy/span, i sum of f0 steps, o f0 count, l/s/_ the same for f1,
 * b sum of f1 steps, w f1 count) */
static void utt_syllable(fao_hist* H, double e, double t, double n, double a, double i, double o, double l, double s,
                         double u_, double b, double w) {
  const double r = 40;
  double T = fa_js_parse_int(e / 2); if (T >= 20) T = 19; hist_inc(&H[H_C], T);
  double k = fa_js_parse_int(t / 2); if (k >= r) k = r - 1; hist_inc(&H[H_U], k);
  double M = fa_js_parse_int(l / 2); if (M >= r) M = r - 1; hist_inc(&H[H_F], M);
  double A = fa_js_parse_int(3 * fa_js_log10(n)); if (A >= 24) A = 23; hist_inc(&H[H_D], A);
  double S = fa_js_parse_int(4 * fa_js_log10(s)); if (S >= 24) S = 23; hist_inc(&H[H_H], S);
  double L = fa_js_parse_int(a / 2); if (L >= 8) L = 7; hist_inc(&H[H_P], L);
  double D = fa_js_parse_int(u_ / 2); if (D >= 8) D = 7; hist_inc(&H[H_M], D);
  double O = fa_js_parse_int(10 * (e - o) / e); if (O >= 10) O = 9; hist_inc(&H[H_G], O);
  double C = fa_js_parse_int(10 * (e - w) / e); if (C >= 10) C = 9; hist_inc(&H[H_Y], C);
  double P = fa_js_parse_int(20 * (i + 50) / 100); if (P >= 20) P = 19; if (P < 0) P = 0; hist_inc(&H[H_V], P);
  double I = fa_js_parse_int(20 * (b + 50) / 100); if (I >= 20) I = 19; if (I < 0) I = 0; hist_inc(&H[H_X], I);
}

/* a() @B107983 over the first n_stores stores; store r is paired with seg_ci[r] (NOT with the segment that produced it:
 * the same misalignment as the time stamps after a dropped segment, quirk 15) */
static void utterance_features(const fao_result* R, int n_stores, double* out /*[264]*/) {
  fao_hist H[H_COUNT];
  memset(H, 0, sizeof(H));
  for (int q = 0; q < H_COUNT; q++) H[q].n = kHistSize[q];
  double prev_end = R->seg_start.d[0];
  for (int r = 0; r < n_stores; r++) {
    const double seg_len = R->seg_len.d[r];
    const int nsyl = R->st_nsyl.d[r];
    double voiced = 0;
    for (int e = 0; e < nsyl; e++) {
      const int sy = R->st_first_syl.d[r] + e;
      const int n = R->syl_len.d[sy];
      const float* F = R->formants + 9 * (size_t)(R->st_row_off.d[r] + R->syl_start.d[sy]);
      double a = 0, i = 0, l = 0, s = 0, c = 0, u = 0, f = 0, d = 0, h = 0, p = 0;
      for (int o = 0; o < n; o++) {
        const float* x = F + 9 * (size_t)o;
        if (x[0] > 0) { c++; a += x[0]; i += x[1]; l += x[2]; if (o > 0) s += (double)x[0] - (double)x[-9]; }
        if (x[3] > 0) { p++; u += x[3]; f += x[4]; d += x[5]; if (o > 0) h += (double)x[3] - (double)x[-9 + 3]; }
      }
      a /= c; i /= c; l /= c; u /= p; f /= p; d /= p;
      utt_syllable(H, n, a, i, l, s, c, u, f, d, h, p);
      voiced += n;
    }
    utt_segment(H, seg_len, nsyl, (double)R->seg_start.d[r] - prev_end, voiced / seg_len);
    prev_end = (double)R->seg_start.d[r] + (double)R->seg_len.d[r];
  }
  for (int q = 0; q < H_COUNT; q++) {        /* w() @B109452 */
    double t = 0;
    for (int k = 0; k < H[q].n; k++) t += H[q].c[k];
    const int norm = !H[q].poison && t > 0;
    for (int k = 0; k < H[q].n; k++) *out++ = norm ? H[q].c[k] / t : H[q].c[k];
  }
}

/* P() @B28869: which seg_ci indices fire a callback, in order.  The stores are indexed by `stored`;
 * the reference indexes seg_ci (u) with the SAME counter, which misaligns after a dropped segment
 * (quirk 15) -- we record the store index and let the host shim reproduce j()/V(). */
static void fire_callbacks(fao_result* R, int* processed) {
  if (R->level == 11) {   /* one callback per P() that sees new stores: b(0, label, clip_time, row264) */
    if (*processed < R->st_len.n) {
      *processed = R->st_len.n;
      double row[264];
      utterance_features(R, R->st_len.n, row);
      for (int q = 0; q < 264; q++) dvec_push(&R->utt_rows, row[q]);
      R->n_utt_rows++;
      ivec_push(&R->cb_si, R->st_len.n - 1);
    }
    return;
  }
  while (*processed < R->st_len.n) {
    int e = (*processed)++;
    int fire = 0;
    if (R->level == 13 || R->level == 12) fire = R->st_nsyl.d[e] > 0;   /* p[e].length > 0 (level 12: unless its first fit threw) */
    else if (R->level == 10) fire = R->st_nsyl.d[e] > 0;
    else if (R->level == 5 || R->level == 4) fire = R->st_len.d[e] > 0;
    else if (R->level == 3) fire = R->st_nsyl.d[e] > 0;   /* s[e].length > 0: at least one ranked track */
    if (fire) ivec_push(&R->cb_si, e);
  }
}

FAO_API fao_result* fao_analyze_frames(const fa_config* c, const uint32_t* frames, int F, int want_trace) {
  fao_result* R = (fao_result*)calloc(1, sizeof(fao_result));
  fao_state st;
  memset(&st, 0, sizeof(st));
  const int B = fa_tab_bands(c);
  const double step = c->window_step_ms, pause = c->pause_length_ms, minlen = c->min_seg_length_ms;
  R->F = F; R->B = B; R->level = c->output_level;
  st.level = c->output_level; st.B = B; st.plot_len = c->plot_len;
  st.max_voiced_bin = (int)fa_js_parse_int(0.7 * (double)B);
  st.window_step = step / 1e3;
  st.seg_breaker = pause > 2 * step ? pause / step : 250 / step;
  st.seg_min_frames = (int)fa_js_parse_int(minlen / step);
  st.auto_gate = c->auto_noise_gate;
  st.c_started = -1;
  if (st.auto_gate) { st.y = 50; st.v = 2; }
  else { st.y = fa_js_pow(10, c->voiced_max_db / 20); st.v = fa_js_pow(10, c->voiced_min_db / 20); }
  st.x = st.y; st.v0 = st.v;
  if ((want_trace & 1) && F > 0) {
    R->tr_n = (int*)calloc(F, sizeof(int)); R->tr_p = (int*)calloc(F, sizeof(int));
    R->tr_cstart = (int*)calloc(F, sizeof(int)); R->tr_cci = (int*)calloc(F, sizeof(int));
    R->tr_nofm = (int*)calloc(F, sizeof(int));
    R->tr_h = (double*)calloc(F, sizeof(double)); R->tr_v = (double*)calloc(F, sizeof(double));
    R->tr_y = (double*)calloc(F, sizeof(double));
  }
  int processed = 0;
  if (c->output_level >= 3) {
    for (int t = 0; t < F; t++) {
      st.current_frame++;
      int fin = process_frame(&st, R, frames + (size_t)t * B, t);
      if (fin != -2) { /* the promise's micro-task: runs before the next frame */
        seg_reset(&st, -1);
        if (fin >= 0) fire_callbacks(R, &processed);
      }
    }
    /* segment_truncate @B30800 (want_trace bit 1: the frames are the prefix of a stream that is still running -- the source has
     * not stopped, nothing is truncated) */
    if (!(want_trace & 2)) {
      int fin = finalize_segment(&st, R, st.c_ci);
      seg_reset(&st, 1);
      if (fin >= 0) fire_callbacks(R, &processed);
    }
  }
  clear_fm(&st);
  free(st.tr);
  R->max_live = st.max_live; R->max_peaks = st.max_peaks;
  return R;
}

FAO_API void fao_free(fao_result* R) {
  if (!R) return;
  free(R->tr_n); free(R->tr_p); free(R->tr_cstart); free(R->tr_cci); free(R->tr_nofm); free(R->tr_h); free(R->tr_v);
  free(R->tr_y); free(R->seg_start.d); free(R->seg_len.d); free(R->seg_stored.d); free(R->st_len.d);
  free(R->st_row_off.d); free(R->st_nsyl.d); free(R->st_first_syl.d); free(R->st_y.d); free(R->st_v.d);
  free(R->st_cs.d); free(R->formants); free(R->energy); free(R->syl_seg.d); free(R->syl_start.d);
  free(R->syl_len.d); free(R->syl_flag.d); free(R->features.d); free(R->cb_si.d); free(R->utt_rows.d);
  free(R->trk_seg.d); free(R->trk_first.d); free(R->trk_n.d); free(R->pt_frame.d); free(R->pt_lo.d); free(R->pt_hi.d);
  free(R->pt_bin.d); free(R->pt_amp.d); free(R->pt_e.d);
  free(R);
}

/* ---- accessors for ctypes ---- */
FAO_API void fao_counts(const fao_result* R, int* out /*[8]*/) {
  out[0] = R->F; out[1] = R->seg_start.n; out[2] = R->st_len.n; out[3] = R->n_rows; out[4] = R->syl_seg.n;
  out[5] = R->n_feature_rows; out[6] = R->cb_si.n; out[7] = R->B;
}
/* level 3: [0] ranked tracks, [1] their points */
FAO_API void fao_track_counts(const fao_result* R, int* out /*[2]*/) { out[0] = R->trk_seg.n; out[1] = R->pt_frame.n; }
FAO_API void fao_get_tracks(const fao_result* R, fa_track* dst) {
  for (int i = 0; i < R->trk_seg.n; i++) {
    dst[i].stored_seg = R->trk_seg.d[i]; dst[i].first_point = R->trk_first.d[i]; dst[i].n_points = R->trk_n.d[i]; dst[i].reserved = 0;
  }
}
FAO_API void fao_get_track_points(const fao_result* R, fa_track_point* dst) {
  for (int i = 0; i < R->pt_frame.n; i++) {
    dst[i].frame = R->pt_frame.d[i]; dst[i].lo = (int16_t)R->pt_lo.d[i]; dst[i].hi = (int16_t)R->pt_hi.d[i];
    dst[i].bin = (int16_t)R->pt_bin.d[i]; dst[i].reserved = 0; dst[i].amp = (uint32_t)R->pt_amp.d[i]; dst[i].energy = R->pt_e.d[i];
  }
}
FAO_API void fao_track_stats(const fao_result* R, int* out /*[2]*/) { out[0] = R->max_live; out[1] = R->max_peaks; }
FAO_API void fao_get_segments(const fao_result* R, fa_segment* dst) {
  for (int i = 0; i < R->seg_start.n; i++) {
    fa_segment* s = &dst[i];
    memset(s, 0, sizeof(*s));
    s->start = R->seg_start.d[i]; s->len = R->seg_len.d[i]; s->stored = R->seg_stored.d[i];
    if (s->stored >= 0) {
      int k = s->stored;
      s->n_syllables = R->st_nsyl.d[k]; s->first_syllable = R->st_first_syl.d[k]; s->row_offset = R->st_row_off.d[k];
      s->ymax = R->st_y.d[k]; s->vmin = R->st_v.d[k]; s->cs_ratio = R->st_cs.d[k];
    } else { s->first_syllable = -1; s->row_offset = -1; }
  }
}
FAO_API void fao_get_formants(const fao_result* R, float* F9, float* E3) {
  if (F9) memcpy(F9, R->formants, sizeof(float) * 9 * (size_t)R->n_rows);
  if (E3) memcpy(E3, R->energy, sizeof(float) * 3 * (size_t)R->n_rows);
}
FAO_API void fao_get_syllables(const fao_result* R, fa_syllable* dst) {
  for (int i = 0; i < R->syl_seg.n; i++) {
    dst[i].stored_seg = R->syl_seg.d[i]; dst[i].start = R->syl_start.d[i]; dst[i].len = R->syl_len.d[i];
    dst[i].reserved = i < R->syl_flag.n ? R->syl_flag.d[i] : 0;
  }
}
FAO_API void fao_get_features(const fao_result* R, double* dst) {
  memcpy(dst, R->features.d, sizeof(double) * (size_t)R->features.n);
}
FAO_API int fao_utterance_rows(const fao_result* R) { return R->n_utt_rows; }
FAO_API void fao_get_utterance_features(const fao_result* R, double* dst) {
  memcpy(dst, R->utt_rows.d, sizeof(double) * (size_t)R->utt_rows.n);
}
FAO_API void fao_get_callbacks(const fao_result* R, int* dst) { memcpy(dst, R->cb_si.d, sizeof(int) * (size_t)R->cb_si.n); }
FAO_API int fao_get_trace(const fao_result* R, int* n, int* p, double* h, double* v, double* y, int* cstart, int* cci,
                          int* nofm) {
  if (!R->tr_n) return -1;
  size_t F = (size_t)R->F;
  memcpy(n, R->tr_n, sizeof(int) * F); memcpy(p, R->tr_p, sizeof(int) * F); memcpy(h, R->tr_h, sizeof(double) * F);
  memcpy(v, R->tr_v, sizeof(double) * F); memcpy(y, R->tr_y, sizeof(double) * F);
  memcpy(cstart, R->tr_cstart, sizeof(int) * F); memcpy(cci, R->tr_cci, sizeof(int) * F);
  memcpy(nofm, R->tr_nofm, sizeof(int) * F);
  return 0;
}

/* ---- whole path on a batch, OpenMP over utterances (bench.py cpu_baseline / --impl reference) ---- */
FAO_API int64_t fao_run_batch(const fa_config* c, const float* pcm, const int64_t* offsets, int n_utt, int sr,
                              int n_threads, int64_t* out_rows /* nullable: feature rows per utterance */) {
  int64_t total_frames = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : total_frames)
  for (int u = 0; u < n_utt; u++) {
    const int64_t n = offsets[u + 1] - offsets[u];
    const int B = fa_tab_bands(c);
    int F = (int)(n / fa_tab_hop(sr, c->window_step_ms));
    uint32_t* frames = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(F > 0 ? F : 1) * B);
    float* spec = c->want_spectrum || c->output_level <= 2
                      ? (float*)malloc(sizeof(float) * (size_t)(F > 0 ? F : 1) * (c->fft_size / 2))
                      : NULL;
    fao_frontend(c, pcm + offsets[u], n, sr, spec, NULL, frames);
    fao_result* R = fao_analyze_frames(c, frames, F, 0);
    if (out_rows) out_rows[u] = R->n_feature_rows;
    total_frames += F;
    fao_free(R);
    free(frames);
    free(spec);
  }
  return total_frames;
}

/* jsmath taps for the literal transliteration and for tests */
FAO_API double fao_js_log10(double x) { return fa_js_log10(x); }
FAO_API double fao_js_log(double x) { return fa_js_log(x); }
FAO_API double fao_js_pow(double x, double y) { return fa_js_pow(x, y); }
