/*
 * fa_curves.h -- output level 12 ("Syllable curves", 23 numbers per syllable): the polynomial fits of make_coeffs
 * (/root/reference/dist/main.js:2@B34527) and polyfit (@B33793), with the parts of numeric@1.2.6 (inner module 5 of the bundle,
 * @B38281; third-party, package.json:9) that they call, restated with IEEE-754 double arithmetic only:
 *
 *   numeric.dotVV / dotMV / dotMMsmall / dotMMbig   @B(38181+9970 / 9625 / 8967 / 9368)   one summation pattern: the LAST
 *        product first, then pairs (x[n] y[n] + x[n-1] y[n-1]) added downwards, then the first product if the length is odd
 *   numeric.inv          Gauss-Jordan, first strictly largest pivot in the column, rows swapped
 *   numeric.gradient     central differences with the step-halving (x /= 16) acceptance test
 *   numeric.uncmin       BFGS on the inverse Hessian with the backtracking line search, tol 1e-8, 1000 iterations
 *   norm2                sqrt of the squares summed from the last element down (mapreduce2 template @B7780)
 *   solve_poly           /root/reference/src/stats.js:3-10   sum of coeffs[c] * Math.pow(x, c), ascending
 *
 * One header for the CPU oracle, and the sm_100a kernel (fa_curves.cu): same operations on the same operands in the same
 * order => the same bits (compile without contraction: -ffp-contract=off / --fmad=false).  Pinned against the reference's own
 * code: oracle/minijs executes the reference's polyfit / make_coeffs and the hand-written numeric functions above
 * (oracle/minijs/numeric_shim.py); tests/golden/ref_js.json holds what they returned.
 *
 * polyfit(e, t, n, log) of the reference:
 *   points r with e[r][t] > 0:  x = r - (first such r),  y = e[r][t]  (or 10 log10 of it),  design row [r^0 .. r^n]  (powers of
 *   the ABSOLUTE row index r, while the objective below uses x: the reference's own inconsistency, kept);
 *   more than two points:  c0 = Float32Array( inv(A^T A) (A^T y) ),  c = uncmin(c -> sum (solve_poly(c, x) - y)^2, c0).solution,
 *   result = [c..., sqrt(objective(c)) / points, points];  otherwise [0 x (n+1), 0, points].
 */
#ifndef FA_CURVES_H_
#define FA_CURVES_H_

#include "fa_jsmath.h"

#ifndef FA_N_CURVE_FEATURES
#define FA_N_CURVE_FEATURES 23 /* /root/reference/src/localstore.js:7 (level 12 -> 23) */
#endif
#define FA_CURVE_MAXD 5        /* coefficients of the largest fit (degree 4) */

/* status of one fit: the reference throws from numeric.uncmin / numeric.gradient in these cases; make_coeffs' try / catch
 * then returns the rows made so far and skips the rest of the segment's syllables */
#define FA_CURVE_OK 0
#define FA_CURVE_THROW_NAN 1       /* "uncmin: f(x0) is a NaN!" / "gradient: f(x) is a NaN!" */
#define FA_CURVE_THROW_GRADIENT 2  /* "Numerical gradient fails" */

typedef struct fa_curve_problem {
  int k;                 /* points */
  int nc;                /* coefficients = degree + 1 */
  const double* prel;    /* [k][nc]  Math.pow(x_p, c) */
  const double* y;       /* [k] */
} fa_curve_problem;

/* numeric.dotVV pattern on strided operands */
FA_HD double fa_num_dot(const double* e, int se, const double* t, int st, int a) {
  double i = e[(a - 1) * se] * t[(a - 1) * st];
  int n;
  for (n = a - 2; n >= 1; n -= 2) i += e[n * se] * t[n * st] + e[(n - 1) * se] * t[(n - 1) * st];
  if (n == 0) i += e[0] * t[0];
  return i;
}

FA_HD int fa_num_isfinite(double x) { return x - x == 0.0; }

/* the objective d(e) of polyfit: sum over the points of (solve_poly(e, x_p) - y_p)^2 */
FA_HD double fa_curve_objective(const fa_curve_problem* P, const double* e) {
  double t = 0;
  for (int p = 0; p < P->k; p++) {
    double ret = 0;
    for (int c = 0; c < P->nc; c++) ret += e[c] * P->prel[p * P->nc + c];
    const double a = ret - P->y[p];
    t += a * a;
  }
  return t;
}

FA_HD double fa_num_max2(double a, double b) { return (a != a || b != b) ? (a != a ? a : b) : (a > b ? a : b); } /* Math.max */
FA_HD double fa_num_min2(double a, double b) { return (a != a || b != b) ? (a != a ? a : b) : (a < b ? a : b); } /* Math.min */
FA_HD double fa_num_abs(double a) { return a < 0 ? -a : (a == 0 ? 0.0 : a); }

/* numeric.gradient(f, x) -> m; returns a FA_CURVE_* status */
FA_HD int fa_num_gradient(const fa_curve_problem* P, const double* t, double* m) {
  const int n = P->nc;
  const double r = fa_curve_objective(P, t);
  if (r != r) return FA_CURVE_THROW_NAN;
  double p[FA_CURVE_MAXD];
  for (int a = 0; a < n; a++) p[a] = t[a];
  int v = 0;
  for (int a = 0; a < n; a++) {
    double x = fa_num_max2(1e-6 * r, 1e-8);
    for (;;) {
      if (++v > 20) return FA_CURVE_THROW_GRADIENT;
      p[a] = t[a] + x;
      const double i = fa_curve_objective(P, p);
      p[a] = t[a] - x;
      const double o = fa_curve_objective(P, p);
      p[a] = t[a];
      if (i != i || o != o) { x /= 16; continue; }
      m[a] = (i - o) / (2 * x);
      const double l = t[a] - x, s = t[a], c = t[a] + x;
      const double u = (i - r) / x, f = (r - o) / x;
      double d = fa_num_max2(fa_num_abs(m[a]), fa_num_abs(r));
      d = fa_num_max2(d, fa_num_abs(i)); d = fa_num_max2(d, fa_num_abs(o)); d = fa_num_max2(d, fa_num_abs(l));
      d = fa_num_max2(d, fa_num_abs(s)); d = fa_num_max2(d, fa_num_abs(c)); d = fa_num_max2(d, 1e-8);
      double w = fa_num_max2(fa_num_abs(u - m[a]), fa_num_abs(f - m[a]));
      w = fa_num_max2(w, fa_num_abs(u - f));
      if (!(fa_num_min2(w / d, x / d) > .001)) break;
      x /= 16;
    }
  }
  return FA_CURVE_OK;
}

/* numeric.uncmin(f, x0): t = x0 in, the solution out.  Returns a FA_CURVE_* status; *iters = N */
FA_HD int fa_num_uncmin(const fa_curve_problem* P, double* t, int* iters) {
  const int u = P->nc;
  const int maxit = 1000;
  double f = fa_curve_objective(P, t);
  if (f != f) return FA_CURVE_THROW_NAN;
  double tol = fa_num_max2(1e-8, 2.220446049250313e-16);
  double k[FA_CURVE_MAXD][FA_CURVE_MAXD];          /* inverse Hessian, identity */
  for (int a = 0; a < u; a++)
    for (int b = 0; b < u; b++) k[a][b] = a == b ? 1.0 : 0.0;
  double m[FA_CURVE_MAXD], p[FA_CURVE_MAXD], y[FA_CURVE_MAXD], v[FA_CURVE_MAXD], g[FA_CURVE_MAXD], x[FA_CURVE_MAXD], h_[FA_CURVE_MAXD];
  int N = 0;
  int rc = fa_num_gradient(P, t, m);
  if (rc != FA_CURVE_OK) return rc;
  while (N < maxit) {
    int ok = 1;
    for (int a = 0; a < u; a++) ok = ok && fa_num_isfinite(m[a]);
    if (!ok) break;                                   /* "Gradient has Infinity or NaN" */
    for (int a = 0; a < u; a++) p[a] = -fa_num_dot(k[a], 1, m, 1, u);   /* neg(dot(Hinv, grad)): dotMV */
    ok = 1;
    for (int a = 0; a < u; a++) ok = ok && fa_num_isfinite(p[a]);
    if (!ok) break;                                   /* "Search direction has Infinity or NaN" */
    double T;
    {
      double acc = 0;
      for (int a = u - 1; a != -1; --a) acc += p[a] * p[a];
      T = fa_sqrt(acc);                               /* norm2 */
    }
    if (T < tol) break;                               /* "Newton step smaller than tol" */
    double w = 1;
    const double c = fa_num_dot(m, 1, p, 1, u);
    double s = f;
    for (int a = 0; a < u; a++) v[a] = t[a];
    /* for(w=1, c, v=t; N<a && !(w*T<n) && (v = add(t, y = mul(p, w)), (s = f(v)) - f >= .1*w*c || isNaN(s)); ) w *= .5, ++N; */
    for (;;) {
      if (!(N < maxit)) break;
      if (w * T < tol) break;
      for (int a = 0; a < u; a++) { y[a] = p[a] * w; v[a] = t[a] + y[a]; }
      s = fa_curve_objective(P, v);
      if (!(s - f >= .1 * w * c || s != s)) break;
      w *= .5;
      ++N;
    }
    if (w * T < tol) break;                           /* "Line search step size smaller than tol" */
    if (N == maxit) break;                            /* "maxit reached during line search" */
    rc = fa_num_gradient(P, v, g);
    if (rc != FA_CURVE_OK) return rc;
    for (int a = 0; a < u; a++) x[a] = g[a] - m[a];
    const double b = fa_num_dot(x, 1, y, 1, u);
    for (int a = 0; a < u; a++) h_[a] = fa_num_dot(k[a], 1, x, 1, u);     /* _ = dot(Hinv, x) */
    /* Hinv = sub(add(Hinv, mul((b + dot(x, _)) / (b*b), tensor(y, y))), div(add(tensor(_, y), tensor(y, _)), b)) */
    const double q = (b + fa_num_dot(x, 1, h_, 1, u)) / (b * b);
    for (int a = 0; a < u; a++)
      for (int bb = 0; bb < u; bb++)
        k[a][bb] = (k[a][bb] + q * (y[a] * y[bb])) - (h_[a] * y[bb] + y[a] * h_[bb]) / b;
    for (int a = 0; a < u; a++) { t[a] = v[a]; m[a] = g[a]; }
    f = s;
    ++N;
  }
  if (iters) *iters = N;
  return FA_CURVE_OK;
}

/* numeric.inv of an n x n matrix (n <= FA_CURVE_MAXD), row-major d -> h */
FA_HD void fa_num_inv(double d[FA_CURVE_MAXD][FA_CURVE_MAXD], double h[FA_CURVE_MAXD][FA_CURVE_MAXD], int n) {
  int rd[FA_CURVE_MAXD], rh[FA_CURVE_MAXD];       /* row permutations (the reference swaps row references) */
  for (int a = 0; a < n; a++) {
    rd[a] = a; rh[a] = a;
    for (int b = 0; b < n; b++) h[a][b] = a == b ? 1.0 : 0.0;
  }
  for (int o = 0; o < n; ++o) {
    int pi = -1;
    double mm = -1;
    for (int i = o; i != n; ++i) {
      const double l = fa_num_abs(d[rd[i]][o]);
      if (l > mm) { pi = i; mm = l; }
    }
    if (pi < 0) pi = o;                              /* a column of NaNs: d[-1] would throw in JS; unreachable for finite input */
    int tswap = rd[pi]; rd[pi] = rd[o]; rd[o] = tswap;
    tswap = rh[pi]; rh[pi] = rh[o]; rh[o] = tswap;
    double* nrow = d[rd[o]];
    double* arow = h[rh[o]];
    const double e = nrow[o];
    for (int l = o; l != n; ++l) nrow[l] /= e;
    for (int l = n - 1; l != -1; --l) arow[l] /= e;
    for (int i = n - 1; i != -1; --i) {
      if (i != o) {
        double* trow = d[rd[i]];
        double* rrow = h[rh[i]];
        const double e2 = trow[o];
        for (int l = o + 1; l != n; ++l) trow[l] -= nrow[l] * e2;
        for (int l = n - 1; l != -1; --l) rrow[l] -= arow[l] * e2;
      }
    }
  }
  /* return the rows in their final order */
  double tmp[FA_CURVE_MAXD][FA_CURVE_MAXD];
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) tmp[a][b] = h[rh[a]][b];
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) h[a][b] = tmp[a][b];
}

/* polyfit @B33793 on column `col` of `len` float32 rows with `stride` floats per row.  work: at least len * (2 * nc + 1)
 * doubles.  out: nc + 2 doubles.  Returns a FA_CURVE_* status (out is only valid for FA_CURVE_OK). */
FA_HD int fa_curve_polyfit(const float* rows, int stride, int col, int len, int degree, int use_log, double* work, double* out) {
  const int nc = degree + 1;
  double* pabs = work;                 /* [k][nc] Math.pow(r, c) */
  double* prel = pabs + (size_t)len * nc;   /* [k][nc] Math.pow(r - u, c) */
  double* y = prel + (size_t)len * nc;      /* [k] */
  int k = 0, u = -1;
  for (int r = 0; r < len; r++) {
    const double val = (double)rows[(size_t)r * stride + col];
    if (val > 0) {
      if (u == -1) u = r;
      y[k] = use_log ? 10 * fa_js_log10(val) : val;
      for (int c = 0; c <= degree; c++) {
        pabs[k * nc + c] = 1 * fa_js_pow((double)r, (double)c);
        prel[k * nc + c] = fa_js_pow((double)(r - u), (double)c);
      }
      k++;
    }
  }
  if (!(k > 2)) {
    for (int c = 0; c < nc; c++) out[c] = 0;
    out[nc] = 0;
    out[nc + 1] = (double)k;
    return FA_CURVE_OK;
  }
  double ata[FA_CURVE_MAXD][FA_CURVE_MAXD], inv[FA_CURVE_MAXD][FA_CURVE_MAXD], aty[FA_CURVE_MAXD], c0[FA_CURVE_MAXD];
  for (int a = 0; a < nc; a++) {
    for (int b = 0; b < nc; b++) ata[a][b] = fa_num_dot(pabs + a, nc, pabs + b, nc, k);   /* dot(transpose(s), s) */
    aty[a] = fa_num_dot(pabs + a, nc, y, 1, k);                                             /* dot(transpose(s), [l]^T) */
  }
  fa_num_inv(ata, inv, nc);
  for (int a = 0; a < nc; a++) c0[a] = (double)(float)fa_num_dot(inv[a], 1, aty, 1, nc);   /* new Float32Array(dot(inv, c)) */
  fa_curve_problem P;
  P.k = k; P.nc = nc; P.prel = prel; P.y = y;
  const int rc = fa_num_uncmin(&P, c0, 0);
  if (rc != FA_CURVE_OK) return rc;
  for (int c = 0; c < nc; c++) out[c] = c0[c];
  out[nc] = fa_sqrt(fa_curve_objective(&P, c0)) / (double)k;
  out[nc + 1] = (double)k;
  return FA_CURVE_OK;
}

/* make_coeffs @B34527 for one syllable: energy rows (3 floats, column 1, degree 4, log) + formant rows (9 floats, columns 0 / 3 /
 * 6, degrees 3 / 3 / 1) -> 23 doubles = [7 | 6 | 6 | 4].  which: 0..3 = one of the four fits (the kernel runs them in four
 * threads); out points at the fit's slice of the row.  work: len * 11 doubles. */
FA_HD int fa_curve_fit_one(const float* F9, const float* E3, int len, int which, double* work, double* out) {
  switch (which) {
    case 0: return fa_curve_polyfit(E3, 3, 1, len, 4, 1, work, out);
    case 1: return fa_curve_polyfit(F9, 9, 0, len, 3, 0, work, out);
    case 2: return fa_curve_polyfit(F9, 9, 3, len, 3, 0, work, out);
    default: return fa_curve_polyfit(F9, 9, 6, len, 1, 0, work, out);
  }
}
FA_HD int fa_curve_slice_offset(int which) { return which == 0 ? 0 : which == 1 ? 7 : which == 2 ? 13 : 19; }

#endif /* FA_CURVES_H_ */
