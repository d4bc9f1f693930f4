/*
 * TEST INFRASTRUCTURE ONLY.  Exhaustive check of the noise gate's threshold map y -> v over EVERY reachable y.
 *
 * C() (/root/reference/dist/main.js:2@B28506) turns the adaptive maximum y into the integer threshold
 *     t = Math.log10(y);  v = t>7 ? parseInt(10^(t-3)/20) : t>6 ? parseInt(10^(t-3)/2) : t>4 ? parseInt(10^(t-2)/2)
 *                           : t>2 ? parseInt(10^(t/3)) : t>1 ? parseInt(y/10) : 1
 * y is always a positive integer below 2^32 (a uint32 amplitude, 2v, or y - parseInt(y/8)), so the map can be enumerated.
 * In exact arithmetic v = floor(y/20000) | floor(y/2000) | floor(y/200) | floor(cbrt y) | floor(y/10) | 1.  Away from the
 * points where that real value IS an integer (multiples of the divisor, perfect cubes) it is at least 1/y >= 2^-32 (relative)
 * from the next integer, thousands of ulps, so ANY log10/pow within a few ulp -- V8's fdlibm port, glibc, this header --
 * gives the same v: engine-independent.  ON those points (224 k of them) the last bit of log10 and pow decides between q
 * and q - 1: this program records which side include/fa_jsmath.h (the fdlibm restatement compiled into oracle, host and
 * kernels) lands on, and tests/test_gate_enumeration.py compares that with a correctly rounded evaluation (mpmath).
 *
 * build + run: gcc -O2 -fopenmp -ffp-contract=off -I include oracle/gate/gate_enum.c -o oracle/_build/gate_enum -lm
 *              oracle/_build/gate_enum > tests/golden/gate_enumeration.json      (about a minute on 8 cores)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fa_jsmath.h"

static double gate_v(double y) {
  const double t = fa_js_log10(y);
  if (t > 7) return fa_js_parse_int(fa_js_pow(10, t - 3) / 20);
  if (t > 6) return fa_js_parse_int(fa_js_pow(10, t - 3) / 2);
  if (t > 4) return fa_js_parse_int(fa_js_pow(10, t - 2) / 2);
  if (t > 2) return fa_js_parse_int(fa_js_pow(10, t / 3));
  if (t > 1) return fa_js_parse_int(y / 10);
  return 1;
}

static uint64_t icbrt(uint64_t y) {
  uint64_t r = 0;
  while ((r + 1) * (r + 1) * (r + 1) <= y) r++;
  return r;
}

/* exact-arithmetic v and whether y sits on an integer point of its branch */
static uint64_t exact_v(uint64_t y, int* on_boundary, int* branch) {
  *on_boundary = 0;
  if (y > 10000000ull) { *branch = 7; *on_boundary = y % 20000 == 0; return y / 20000; }
  if (y > 1000000ull) { *branch = 6; *on_boundary = y % 2000 == 0; return y / 2000; }
  if (y > 10000ull) { *branch = 4; *on_boundary = y % 200 == 0; return y / 200; }
  if (y > 100ull) { *branch = 2; const uint64_t r = icbrt(y); *on_boundary = r * r * r == y; return r; }
  if (y > 10ull) { *branch = 1; return y / 10; }
  *branch = 0;
  return 1;
}

int main(int argc, char** argv) {
  const uint64_t hi = argc > 1 ? strtoull(argv[1], 0, 0) : (1ull << 32);
  uint64_t n_checked = 0, n_off_boundary_bad = 0, n_boundary = 0, n_boundary_low = 0, n_other = 0;
  uint64_t first_bad = 0;
  /* boundary points where the header lands BELOW the exact integer (v = q - 1), per branch, in order */
  enum { CAP = 1 << 20 };
  uint32_t* low = (uint32_t*)malloc(sizeof(uint32_t) * CAP);
  uint64_t n_low_list = 0;
  uint64_t per_branch[8][2];
  memset(per_branch, 0, sizeof(per_branch));
#pragma omp parallel for schedule(dynamic, 1 << 20) reduction(+ : n_checked, n_off_boundary_bad, n_boundary, n_boundary_low, n_other)
  for (uint64_t y = 1; y < hi; y++) {
    int ob, br;
    const uint64_t q = exact_v(y, &ob, &br);
    const double v = gate_v((double)y);
    n_checked++;
    if (!ob) {
      if (v != (double)q) {
        n_off_boundary_bad++;
#pragma omp critical
        if (!first_bad || y < first_bad) first_bad = y;
      }
    } else {
      n_boundary++;
#pragma omp atomic
      per_branch[br][0]++;
      if (v == (double)q) {
      } else if (v == (double)q - 1) {
        n_boundary_low++;
#pragma omp atomic
        per_branch[br][1]++;
#pragma omp critical
        { if (n_low_list < CAP) low[n_low_list++] = (uint32_t)y; }
      } else n_other++;
    }
  }
  /* sort the list (threads append out of order) */
  for (uint64_t i = 1; i < n_low_list; i++) { uint32_t k = low[i]; uint64_t j = i; while (j > 0 && low[j - 1] > k) { low[j] = low[j - 1]; j--; } low[j] = k; }
  printf("{\n \"range\": [1, %llu],\n \"checked\": %llu,\n \"off_boundary_mismatches\": %llu,\n \"first_off_boundary_mismatch\": %llu,\n",
         (unsigned long long)hi, (unsigned long long)n_checked, (unsigned long long)n_off_boundary_bad, (unsigned long long)first_bad);
  printf(" \"boundary_points\": %llu,\n \"boundary_points_low\": %llu,\n \"boundary_points_neither_q_nor_q_minus_1\": %llu,\n",
         (unsigned long long)n_boundary, (unsigned long long)n_boundary_low, (unsigned long long)n_other);
  printf(" \"per_branch\": {");
  const int brs[4] = {2, 4, 6, 7};
  for (int i = 0; i < 4; i++)
    printf("%s\"t>%d\": {\"boundary_points\": %llu, \"low\": %llu}", i ? ", " : "", brs[i], (unsigned long long)per_branch[brs[i]][0],
           (unsigned long long)per_branch[brs[i]][1]);
  printf("},\n \"low_points\": [");
  for (uint64_t i = 0; i < n_low_list; i++) printf("%s%u", i ? "," : "", low[i]);
  printf("]\n}\n");
  return 0;
}
