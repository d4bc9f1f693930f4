// fa_curves.cu -- K8: output level 12 ("Syllable curves"), 23 doubles per syllable.
//
// make_coeffs (/root/reference/dist/main.js:2@B34527) fits four polynomials per syllable with polyfit (@B33793): the energy
// column in dB (degree 4) and the formant columns 0 / 3 / 6 (degrees 3 / 3 / 1); each fit is a normal-equation solve
// (numeric.inv) rounded to float32 and then refined by numeric.uncmin (BFGS with numerical gradients, inner module 5 @B38281).
// The arithmetic lives in include/fa_curves.h, the SAME header the CPU oracle compiles, so the doubles are bit-identical
// (FP64, no contraction).
//
// Mapping: a fit is a strictly sequential FP64 iteration (every objective value feeds the next line-search test); fits are
// independent.  One THREAD per (syllable, fit), and the threads of a warp take 32 DIFFERENT syllables of the same fit kind:
//   K8a fa_curves_list_kernel   flattens the sub-batch's syllables into a work list (utterance, syllable) -- order is free,
//                               results land at the syllable's own row;
//   K8b fa_curves_kernel        grid.y = the four fit kinds (same degree per warp => same loop bounds), thread = list entry;
//   K8c fa_curves_flag_kernel   the reference's try / catch rule: a fit that throws inside numeric (NaN objective, failing
//                               numerical gradient) ends the segment's row list -- that syllable and the later ones of the segment
//                               get a NaN row and a flag.
// Measured on C2 (3030 syllables): lanes = the syllables of ONE utterance (3 of 32 lanes busy) 5.5 ms; a warp per fit with the
// objective's points spread over the lanes 11.3 ms (the FP64 pipe, not latency, is the limit: 32 lanes repeating the scalar
// algebra); this mapping: see DESIGN.md.  The powers of the abscissae and the ordinates of a fit sit in a global work area (the
// syllable's own rows, 34 doubles per row), re-read from L1/L2 at every evaluation of the objective.
#include "fa_curves.h"
#include "fa_internal.cuh"

namespace {

constexpr int kCurveThreads = 128;
constexpr int kWorkPerRow = 34;      // doubles of work area per formant row: 11 + 9 + 9 + 5

__global__ void __launch_bounds__(kCurveThreads) fa_curves_list_kernel(const FaCurveParams p) {
  const int ui = blockIdx.x * kCurveThreads + threadIdx.x;
  if (ui >= p.utt_count) return;
  const int u = p.utt_begin + ui;
  const int n = p.n_syls[u];
  p.n_feat[u] = n;
  if (n == 0) return;
  const int base = atomicAdd(p.list_count, n);
  for (int i = 0; i < n; i++) p.list[base + i] = make_int2(u, i);
}

__global__ void __launch_bounds__(kCurveThreads, 2) fa_curves_kernel(const FaCurveParams p) {
  const int idx = blockIdx.x * kCurveThreads + threadIdx.x;
  if (idx >= *p.list_count) return;
  const int2 it = p.list[idx];
  const int u = it.x, row = it.y, which = blockIdx.y;
  const long long row0 = p.frame_off[u], sb = row0 + u;
  const fa_syllable sy = p.syls[sb + row];
  int s = 0;
  while (p.segs[sb + s].stored != sy.stored_seg) s++;   // stored indices increase in seg_ci order
  const size_t r = (size_t)(row0 + (p.epochs ? p.epochs[sb + s].first : p.segs[sb + s].row_offset) + sy.start);
  const int len = sy.len;
  const int woff = which == 0 ? 0 : which == 1 ? 11 : which == 2 ? 20 : 29;
  double* work = p.work + r * kWorkPerRow + (size_t)len * woff;
  double* out = p.rows + (size_t)(sb + row) * FA_N_CURVE_FEATURES + fa_curve_slice_offset(which);
  p.status[(size_t)(sb + row) * 4 + which] = fa_curve_fit_one(p.formants + r * 9, p.energy + r * 3, len, which, work, out);
}

__global__ void __launch_bounds__(kCurveThreads) fa_curves_flag_kernel(const FaCurveParams p) {
  const int u = p.utt_begin + blockIdx.x;
  const long long row0 = p.frame_off[u], sb = row0 + u;
  const int R = p.n_syls[u];
  for (int row = threadIdx.x; row < R; row += kCurveThreads) {
    const fa_syllable sy = p.syls[sb + row];
    // syllables of one stored segment are consecutive: walk back to the segment's first one
    int thrown = 0;
    for (int q = row; q >= 0 && p.syls[sb + q].stored_seg == sy.stored_seg; q--) {
      const int* st = p.status + (size_t)(sb + q) * 4;
      thrown |= st[0] | st[1] | st[2] | st[3];
    }
    if (thrown) {
      p.syls[sb + row].reserved = 1;
      double* out = p.rows + (size_t)(sb + row) * FA_N_CURVE_FEATURES;
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      for (int q = 0; q < FA_N_CURVE_FEATURES; q++) out[q] = nan;
    }
  }
}

}  // namespace

cudaError_t fa_launch_curves(const FaCurveParams& p, cudaStream_t s, int* launches) {
  if (p.utt_count <= 0) return cudaSuccess;
  fa_curves_list_kernel<<<(p.utt_count + kCurveThreads - 1) / kCurveThreads, kCurveThreads, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // a syllable is at least two frames long and followed by a quiet frame: list_cap bounds the list from above
  fa_curves_kernel<<<dim3((unsigned)((p.list_cap + kCurveThreads - 1) / kCurveThreads), 4), kCurveThreads, 0, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fa_curves_flag_kernel<<<p.utt_count, kCurveThreads, 0, s>>>(p);
  if (launches) (*launches) += 3;
  return cudaGetLastError();
}
