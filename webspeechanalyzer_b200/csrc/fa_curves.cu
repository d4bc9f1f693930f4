// fa_curves.cu -- K8: output level 12 ("Syllable curves"), 23 doubles per syllable.
//
// make_coeffs (/root/reference/dist/main.js:2@B34527) fits four polynomials per syllable with polyfit (@B33793): the energy
// column in dB (degree 4) and the formant columns 0 / 3 / 6 (degrees 3 / 3 / 1); each fit is a normal-equation solve
// (numeric.inv) rounded to float32 and then refined by numeric.uncmin (BFGS with numerical gradients, inner module 5 @B38281).
// The arithmetic lives in include/fa_curves.h, the SAME header the CPU oracle compiles, so the doubles are bit-identical
// (FP64, no contraction).
//
// Mapping: the work per fit is a strictly sequential FP64 iteration (every objective value feeds the next line-search test), a few
// hundred to a few thousand dependent operations long, on kilobytes of data -- latency bound by nature.  One THREAD per
// (syllable, fit); the four warps of a CTA take the four fit kinds (same degree per warp => no divergence on the loop bounds),
// lanes take the utterance's syllables.  The powers of the abscissae and the ordinates of a fit sit in a global work area (its
// slice = the syllable's own rows, 34 doubles per row), where they are re-read from L1/L2 at every evaluation of the objective.
// A second tiny kernel applies the reference's try / catch rule: a fit that throws inside numeric (NaN objective, failing
// numerical gradient) ends the segment's row list -- that syllable and the later ones of the segment get a NaN row and a flag.
#include "fa_curves.h"
#include "fa_internal.cuh"

namespace {

constexpr int kCurveThreads = 128;   // 4 warps = 4 fit kinds
constexpr int kWorkPerRow = 34;      // doubles of work area per formant row: 11 + 9 + 9 + 5

__device__ __forceinline__ int seg_of_syllable(const FaCurveParams& p, long long sb, const fa_syllable& sy) {
  int s = 0;
  while (p.segs[sb + s].stored != sy.stored_seg) s++;   // stored indices increase in seg_ci order
  return s;
}

__global__ void __launch_bounds__(kCurveThreads) fa_curves_kernel(const FaCurveParams p) {
  const int u = p.utt_begin + blockIdx.x;
  const long long row0 = p.frame_off[u], sb = row0 + u;
  const int R = p.n_syls[u];
  if (threadIdx.x == 0 && blockIdx.y == 0) p.n_feat[u] = R;
  const int which = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = lane + 32 * (int)blockIdx.y; row < R; row += 32 * (int)gridDim.y) {
    const fa_syllable sy = p.syls[sb + row];
    const int s = seg_of_syllable(p, sb, sy);
    const size_t r = (size_t)(row0 + (p.epochs ? p.epochs[sb + s].first : p.segs[sb + s].row_offset) + sy.start);
    const int len = sy.len;
    const int woff = which == 0 ? 0 : which == 1 ? 11 : which == 2 ? 20 : 29;
    double* work = p.work + r * kWorkPerRow + (size_t)len * woff;
    double* out = p.rows + (size_t)(sb + row) * FA_N_CURVE_FEATURES + fa_curve_slice_offset(which);
    p.status[(size_t)(sb + row) * 4 + which] = fa_curve_fit_one(p.formants + r * 9, p.energy + r * 3, len, which, work, out);
  }
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
  fa_curves_kernel<<<dim3(p.utt_count, p.row_slices > 0 ? p.row_slices : 1), kCurveThreads, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fa_curves_flag_kernel<<<p.utt_count, kCurveThreads, 0, s>>>(p);
  if (launches) (*launches) += 2;
  return cudaGetLastError();
}
