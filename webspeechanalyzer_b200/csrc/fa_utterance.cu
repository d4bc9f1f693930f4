// fa_utterance.cu -- K6: the 264-dim utterance distributions of output level 11.
//
// Restates get_utterance_features (/root/reference/dist/main.js:2@B107983, inner module 7 of the formantanalyzer
// bundle; per-syllable binning b() @B109255, per-segment binning _() @B109020, normalisation w() @B109452): 15 histograms
// (sizes 10,10,10,10,20,40,40,24,24,8,8,10,10,20,20 = 264 bins) over the syllable tables K3 produced.  The reference calls
// it from P() @B28869 every time a segment has been stored and hands the callback the distribution over ALL stores so
// far, so the output is one cumulative row per stored segment.
//
// Mapping: one warp per utterance.  Stores are visited in order; the syllables of a store are spread over the lanes, each
// lane reduces its syllable's Float32Array(9) rows (sums of integer-valued doubles: exact, order-free) and bumps integer
// bin counters in shared memory with atomics; then the 32 lanes write the row, normalised per histogram.  Integer counts
// and one IEEE division per bin: bit-identical to the reference's doubles.
//
// Reference quirks kept: (1) `hist[k]++` with k = NaN or k <= -1 creates a named property holding NaN instead of touching
// a bin; w()'s for-in sum then is NaN, `t > 0` fails and that histogram stays UN-normalised (a "poison" bit per histogram
// here); (2) store r is paired with seg_ci[r], not with the segment that produced it -- after a dropped segment the
// segment-level bins (length, gap) use the wrong entry, exactly like the time stamps.
#include "fa_internal.cuh"
#include "fa_jsmath.h"

namespace {

constexpr int kUttWarps = 4;
constexpr int kBins = FA_N_UTT_FEATURES;   // 264
constexpr int kHists = 15;
// first bin of histogram q (module-level arrays i,o,l,s,c,u,f,d,h,p,m,g,y,v,x in the reference's output order)
__constant__ int c_hist_off[kHists + 1] = {0, 10, 20, 30, 40, 60, 100, 140, 164, 188, 196, 204, 214, 224, 244, 264};

enum { H_I, H_O, H_L, H_S, H_C, H_U, H_F, H_D, H_H, H_P, H_M, H_G, H_Y, H_V, H_X };

// arr[k]++ of the reference for a double index k that has been clamped from above already
__device__ __forceinline__ void bump(int* hist, unsigned* poison, const int q, const double k) {
  if (k != k || k <= -1.0) { atomicOr(poison, 1u << q); return; }   // named property "NaN" / "-n": poisons w()'s sum
  atomicAdd(&hist[c_hist_off[q] + (int)k], 1);                      // (-1, 0) truncates to -0 -> bin 0
}

__device__ __forceinline__ double clamp_hi(double x, const double n) { return x >= n ? n - 1 : x; }
__device__ __forceinline__ double clamp_both(double x, const double n) { x = x >= n ? n - 1 : x; return x < 0 ? 0 : x; }

__global__ void __launch_bounds__(kUttWarps * 32) fa_utterance_kernel(const FaUtteranceParams p) {
  __shared__ int s_hist[kUttWarps][kBins];
  __shared__ unsigned s_poison[kUttWarps];
  __shared__ double s_tot[kUttWarps][kHists + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ui = blockIdx.x * kUttWarps + warp;
  if (ui >= p.utt_count) return;
  const int u = p.utt_begin + ui;
  const long long row0 = p.frame_off[u], sb = row0 + u;
  const int nseg = p.n_segs[u];
  const long long base = p.row_base[u];
  const int cap = (int)(p.row_base[u + 1] - base);
  int* hist = s_hist[warp];
  for (int j = lane; j < kBins; j += 32) hist[j] = 0;
  if (lane == 0) s_poison[warp] = 0;
  __syncwarp();
  int k = 0;                // stores so far
  double prev_end = nseg > 0 ? (double)p.segs[sb].start : 0.0;
  for (int s = 0; s < nseg; s++) {
    const fa_segment sg = p.segs[sb + s];
    if (sg.stored < 0) continue;          // dropped by the throw in straighten_formants: seg_ci keeps it, the stores do not
    const fa_segment pair = p.segs[sb + k];   // e[r] of a(): seg_ci[store index]
    int voiced = 0;
    for (int e = lane; e < sg.n_syllables; e += 32) {
      const fa_syllable sy = p.syls[sb + sg.first_syllable + e];
      const float* F = p.formants + (size_t)(row0 + (p.epochs ? p.epochs[sb + s].first : sg.row_offset) + sy.start) * 9;
      double a = 0, en = 0, sp = 0, st = 0, c = 0, a1 = 0, en1 = 0, sp1 = 0, st1 = 0, c1 = 0;
      float prev0 = 0.f, prev3 = 0.f;
      for (int o = 0; o < sy.len; o++) {
        const float x0 = F[o * 9 + 0], x3 = F[o * 9 + 3];
        if (x0 > 0.f) { c += 1; a += (double)x0; en += (double)F[o * 9 + 1]; sp += (double)F[o * 9 + 2]; if (o > 0) st += (double)x0 - (double)prev0; }
        if (x3 > 0.f) { c1 += 1; a1 += (double)x3; en1 += (double)F[o * 9 + 4]; sp1 += (double)F[o * 9 + 5]; if (o > 0) st1 += (double)x3 - (double)prev3; }
        prev0 = x0; prev3 = x3;
      }
      a /= c; en /= c; sp /= c; a1 /= c1; en1 /= c1; sp1 /= c1;
      const double len = (double)sy.len;
      unsigned* po = &s_poison[warp];
      bump(hist, po, H_C, clamp_hi(fa_js_parse_int(len / 2), 20));
      bump(hist, po, H_U, clamp_hi(fa_js_parse_int(a / 2), 40));
      bump(hist, po, H_F, clamp_hi(fa_js_parse_int(a1 / 2), 40));
      bump(hist, po, H_D, clamp_hi(fa_js_parse_int(3 * fa_js_log10(en)), 24));
      bump(hist, po, H_H, clamp_hi(fa_js_parse_int(4 * fa_js_log10(en1)), 24));
      bump(hist, po, H_P, clamp_hi(fa_js_parse_int(sp / 2), 8));
      bump(hist, po, H_M, clamp_hi(fa_js_parse_int(sp1 / 2), 8));
      bump(hist, po, H_G, clamp_hi(fa_js_parse_int(10 * (len - c) / len), 10));
      bump(hist, po, H_Y, clamp_hi(fa_js_parse_int(10 * (len - c1) / len), 10));
      bump(hist, po, H_V, clamp_both(fa_js_parse_int(20 * (st + 50) / 100), 20));
      bump(hist, po, H_X, clamp_both(fa_js_parse_int(20 * (st1 + 50) / 100), 20));
      voiced += sy.len;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) voiced += __shfl_xor_sync(0xffffffffu, voiced, d);
    if (lane == 0) {
      unsigned* po = &s_poison[warp];
      const double seg_len = (double)pair.len;
      bump(hist, po, H_I, clamp_hi(fa_js_parse_int(10 * seg_len / 150), 10));
      bump(hist, po, H_O, clamp_hi((double)sg.n_syllables, 10));
      bump(hist, po, H_L, clamp_hi(fa_js_parse_int(10 * ((double)pair.start - prev_end) / 150), 10));
      bump(hist, po, H_S, clamp_both(fa_js_parse_int(2 * ((double)voiced / seg_len - .3) * 10), 10));
    }
    prev_end = (double)pair.start + (double)pair.len;
    __syncwarp();
    // w(): per-histogram totals, then the row of this store
    if (lane < kHists) {
      int t = 0;
      for (int j = c_hist_off[lane]; j < c_hist_off[lane + 1]; j++) t += hist[j];
      s_tot[warp][lane] = (double)t;
    }
    __syncwarp();
    if (k < cap) {
      const unsigned poison = s_poison[warp];
      double* out = p.rows + (size_t)(base + k) * kBins;
      for (int j = lane; j < kBins; j += 32) {
        int q = 0;
        while (j >= c_hist_off[q + 1]) q++;
        const double t = s_tot[warp][q], v = (double)hist[j];
        out[j] = (!((poison >> q) & 1u) && t > 0) ? v / t : v;
      }
    } else if (lane == 0) {
      p.overflow[u] = 1;
    }
    __syncwarp();
    k++;
  }
  if (lane == 0) p.n_feat[u] = k < cap ? k : cap;
}

}  // namespace

cudaError_t fa_launch_utterance(const FaUtteranceParams& p, cudaStream_t s, int* launches) {
  if (p.utt_count <= 0) return cudaSuccess;
  fa_utterance_kernel<<<(p.utt_count + kUttWarps - 1) / kUttWarps, kUttWarps * 32, 0, s>>>(p);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
