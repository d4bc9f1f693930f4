// fa_peaks.cu -- K2: per-frame candidate peak scan (stage S2).
//
// Restates the scan loop of D() (/root/reference/dist/main.js:2@B25863): a 3-neighbour strict
// rise/fall automaton over bins 1..B-1 that closes a peak [lo, hi, pk] on the next rise after a
// fall, after three "flat" bins, or at the last bin, with lo/hi trimmed to bins >= e[pk]/10.
// The adaptive threshold v only gates whether close() RECORDS a peak, so this
// kernel emits every candidate (frame-parallel, v-independent) and K3 applies e[pk] > v inside its
// sequential scan.  Also emits g = sum e[1..B-1] (exact: < 2^53).
//
// Mapping: the automaton is serial in the bin index, frames are independent -> one thread per
// frame; a CTA stages 64 frames (64 x B uint32) in shared memory with coalesced loads, rows padded
// by one word so the per-thread row walks are bank-conflict free.  Integer compares only.
#include "fa_internal.cuh"

namespace {

constexpr int kFramesPerCta = 64;

__global__ void __launch_bounds__(kFramesPerCta) fa_peaks_kernel(const FaPeaksParams p) {
  extern __shared__ uint32_t s_e[];  // [64][B + 1]
  const int B = p.B, ld = B + 1;
  const long long f0 = (long long)blockIdx.x * kFramesPerCta;
  const int nf = (int)min((long long)kFramesPerCta, p.n_frames - f0);
  const uint32_t* __restrict__ src = p.frames + (size_t)f0 * B;
  for (int i = threadIdx.x; i < nf * B; i += kFramesPerCta) {
    const int r = i / B, c = i - r * B;
    s_e[r * ld + c] = __ldg(src + i);
  }
  __syncthreads();
  if ((int)threadIdx.x >= nf) return;
  const uint32_t* e = s_e + threadIdx.x * ld;
  uint32_t* out = p.cand + (size_t)(f0 + threadIdx.x) * p.maxp;
  int n = 0, lo = 0, pk = 0, hi = 0, flat = 0, dir = 0;
  unsigned long long g = 0;

  auto emit = [&](int last) {
    int l2 = lo, h2 = hi;
    // e[i] < e[pk]/10 in doubles  <=>  10*e[i] < e[pk] in integers (both exact)
    const unsigned long long top = e[pk];
    while (l2 < pk && 10ull * e[l2] < top) l2++;
    while (h2 > pk && 10ull * e[h2] < top) h2--;
    if (n < p.maxp) out[n] = (uint32_t)l2 | ((uint32_t)h2 << 8) | ((uint32_t)pk << 16) | ((uint32_t)last << 24);
    n++;
  };

  uint32_t e1 = e[0], e2 = 0, e3 = 0;  // e[a-1], e[a-2], e[a-3]
  for (int a = 1; a < B; a++) {
    const uint32_t ea = e[a];
    g += ea;
    const bool rise = ea > e1 && (a < 2 || ea > e2) && (a < 3 || ea > e3);
    const bool fall = ea < e1 && (a < 2 || ea < e2) && (a < 3 || ea < e3);
    if (rise) {
      if (dir != 1) {
        if (dir == -1 && lo <= pk && pk < hi) emit(0);
        lo = a - 1;
        pk = a;
      } else {
        pk = a;
      }
      dir = 1;
    } else if (fall) {
      if (dir != 0) { hi = a; dir = -1; }
    } else if (dir == -1) {
      if (++flat > 2) {
        flat = 0;
        if (lo <= pk && pk < hi) emit(0);
        dir = 0;
      }
    } else if (dir == 1 && ea > e1) {
      pk = a;
    }
    if (a == B - 1 && dir == 1) {
      hi = a;
      pk = a;
      if (lo < pk && pk <= hi) emit(1);
    }
    e3 = e2; e2 = e1; e1 = ea;
  }
  p.ncand[f0 + threadIdx.x] = n;
  p.gsum[f0 + threadIdx.x] = (double)g;
}

}  // namespace

cudaError_t fa_launch_peaks(const FaPeaksParams& p, cudaStream_t s, int* launches) {
  if (p.n_frames <= 0) return cudaSuccess;
  const int bytes = kFramesPerCta * (p.B + 1) * 4;
  cudaError_t e = cudaFuncSetAttribute(fa_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  const long long grid = (p.n_frames + kFramesPerCta - 1) / kFramesPerCta;
  fa_peaks_kernel<<<(unsigned)grid, kFramesPerCta, bytes, s>>>(p);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
