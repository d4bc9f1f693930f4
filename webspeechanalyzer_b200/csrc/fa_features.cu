// fa_features.cu -- K4: the 53 statistics per segment (level 5) or per syllable (level 13), and
// K5: gather of the per-utterance tables into dense arrays.
//
// Restates formant_features (/root/reference/dist/main.js:2@B32369, vector assembly @B33436) and the
// stats helpers it calls (/root/reference/src/stats.js:29-64: array_mean_NZ, only_std_NZ,
// mean_std_NZ, arraySum).  A row is a segmented reduction over its frames; the three formant slots
// of a row are independent, so the unit of work is (row, slot): one thread walks the slot's frames
// in array order (the reference's summation order, so the doubles come out bit-identical to the
// oracle's) in two passes -- sums and the accent automaton first, squared deviations second.
// FP64 throughout, log10 from include/fa_jsmath.h.  Kilobytes per row: latency bound, not HBM bound.
#include "fa_internal.cuh"
#include "fa_jsmath.h"

namespace {

constexpr int kFeatThreads = 64;

__device__ void slot_features(const float* __restrict__ F, const int len, const int n, const double ymax,
                              double* __restrict__ out /*16*/) {
  double cnt = 0, runs = 0, up = 0, down = 0;
  double sum_c = 0, sum_w = 0, sum_T = 0, sum_k = 0, sum_knz = 0, sum_M = 0, sum_Anz = 0;
  int m = 0, n_knz = 0, na = 0, n_anz = 0;
  {
    bool prev = false;
    int S = 0;
    double L = 0;
    for (int t = 0; t < len; t++) {
      const double r = (double)F[(size_t)t * 9 + 3 * n], a = (double)F[(size_t)t * 9 + 3 * n + 1];
      if (r > 0 && a > 0) {
        const double f = (double)F[(size_t)t * 9 + 3 * n + 2], d = 20.0 * fa_js_log10(a);
        sum_c += r * d; sum_w += r; sum_M += f * d; sum_T += a; sum_k += d;
        if (d > 0) { sum_knz += d; n_knz++; }
        m++;
        if (prev) {
          const double j = r - (double)F[(size_t)(t - 1) * 9 + 3 * n];
          if (j > 1) up += j; else if (j < -1) down += -1 * j;
          if (a > L) { L = a; S = 1; }
          else if (S == 1 && a < L / 2) {
            if (L > 10) { na++; if (d > 0) { sum_Anz += d; n_anz++; } }
            L = 0; S = -1;
          }
        }
        if (!prev) runs += 1;
        prev = true;
        cnt += 1;
      } else { prev = false; S = 0; L = 0; }
    }
  }
  double fmean = 0, fstd = 0, dbm = 0, dbs = 0, e_len = 0, e_cnt = 0, span = 0, accm = 0, accs = 0, prom = 0;
  if (runs > 0) {
    e_len = sum_T / (double)len * 100 / ymax;
    e_cnt = sum_T / cnt * 100 / ymax;
    fmean = sum_c / sum_k;
    span = sum_M / sum_k;
    const double mean_w = sum_w / (double)m;  // every bin in the list is > 0
    dbm = sum_knz / (double)n_knz;
    if (na > 0) accm = sum_Anz / (double)n_anz;
    double acc_w = 0, acc_k = 0, acc_a = 0;
    bool prev = false;
    int S = 0;
    double L = 0;
    for (int t = 0; t < len; t++) {
      const double r = (double)F[(size_t)t * 9 + 3 * n], a = (double)F[(size_t)t * 9 + 3 * n + 1];
      if (r > 0 && a > 0) {
        const double d = 20.0 * fa_js_log10(a);
        const double dw = r - mean_w, dk = d - dbm;
        acc_w += dw * dw;
        acc_k += dk * dk;
        if (prev) {
          if (a > L) { L = a; S = 1; }
          else if (S == 1 && a < L / 2) {
            if (L > 10) { const double da = d - accm; acc_a += da * da; }
            L = 0; S = -1;
          }
        }
        prev = true;
      } else { prev = false; S = 0; L = 0; }
    }
    fstd = fa_sqrt(acc_w / (double)m);
    dbs = fa_sqrt(acc_k / (double)m);
    if (na > 0) {
      accs = fa_sqrt(acc_a / (double)na);
      prom = 100 * (accm / (sum_k / (double)m) - 1);
    }
  }
  out[0] = fmean; out[1] = fstd; out[2] = dbm; out[3] = dbs; out[4] = e_len; out[5] = e_cnt; out[6] = span;
  out[7] = cnt; out[8] = runs; out[9] = up; out[10] = down; out[11] = (double)na; out[12] = accm; out[13] = accs;
  out[14] = prom; out[15] = 100 * cnt / (double)len;
}

__global__ void __launch_bounds__(kFeatThreads) fa_features_kernel(const FaFeatureParams p) {
  const int u = p.utt_begin + blockIdx.x;
  const long long row0 = p.frame_off[u], sb = row0 + u;
  const int nseg = p.n_segs[u];
  const bool per_syl = p.level == 13;
  const int nrows = per_syl ? p.n_syls[u] : 0;
  // level 5: rows are the stored segments, in seg_ci order
  __shared__ int s_rows;
  if (!per_syl) {
    if (threadIdx.x == 0) {
      int c = 0;
      for (int s = 0; s < nseg; s++) c += p.segs[sb + s].stored >= 0;
      s_rows = c;
    }
    __syncthreads();
  }
  const int R = per_syl ? nrows : s_rows;
  if (threadIdx.x == 0) p.n_feat[u] = R;
  for (int item = threadIdx.x; item < R * 3; item += kFeatThreads) {
    const int row = item / 3, slot = item - row * 3;
    const float* F;
    int len;
    const fa_segment* sg;
    if (per_syl) {
      const fa_syllable sy = p.syls[sb + row];
      // find the owning segment: stored indices are increasing in seg_ci order
      int s = 0;
      while (p.segs[sb + s].stored != sy.stored_seg) s++;
      sg = &p.segs[sb + s];
      F = p.formants + (size_t)(row0 + sg->row_offset + sy.start) * 9;
      len = sy.len;
    } else {
      int s = 0, c = -1;
      for (;; s++) { if (p.segs[sb + s].stored >= 0 && ++c == row) break; }
      sg = &p.segs[sb + s];
      F = p.formants + (size_t)(row0 + sg->row_offset) * 9;
      len = sg->len;
    }
    double* out = p.features + (size_t)(sb + row) * FA_N_FEATURES;
    if (slot == 0) {
      out[0] = (double)len;
      out[1] = fa_sqrt((double)len);
      out[2] = sg->cs_ratio;
      out[3] = fa_js_log10(sg->ymax);
      out[4] = sg->vmin;
    }
    slot_features(F, len, slot, sg->ymax, out + 5 + 16 * slot);
  }
}

// ---- K5: dense gather ----
__global__ void __launch_bounds__(1024) fa_prefix_kernel(const FaGatherArgs g) {
  __shared__ long long s_part[4][32];
  __shared__ long long s_run[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 4) s_run[tid] = 0;
  __syncthreads();
  const int* src[4] = {g.n_segs, g.n_rows, g.n_syls, g.n_feat};
  for (int base = 0; base < g.n_utt; base += 1024) {
    const int u = base + tid;
    long long v[4], incl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      v[k] = u < g.n_utt ? src[k][u] : 0;
      long long x = v[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
      incl[k] = x;
      if (lane == 31) s_part[k][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        long long x = s_part[k][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const long long t = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += t;
        }
        s_part[k][lane] = x;  // inclusive over warps
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long before = s_run[k] + (warp ? s_part[k][warp - 1] : 0) + incl[k] - v[k];
      if (u < g.n_utt) g.off[(size_t)k * (g.n_utt + 1) + u] = before;
    }
    __syncthreads();
    if (tid < 4) s_run[tid] += s_part[tid][31];
    __syncthreads();
  }
  if (tid < 4) g.off[(size_t)tid * (g.n_utt + 1) + g.n_utt] = s_run[tid];
}

__global__ void __launch_bounds__(128) fa_gather_kernel(const FaGatherArgs g) {
  const int u = blockIdx.x, tid = threadIdx.x;
  const long long row0 = g.frame_off[u], sb = row0 + u;
  const long long* off = g.off;
  const int N1 = g.n_utt + 1;
  {
    const int n = g.n_segs[u];
    const long long o = off[u];
    const int words = n * (int)(sizeof(fa_segment) / 4);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(g.segs + sb);
    uint32_t* d = reinterpret_cast<uint32_t*>(g.d_segs + o);
    for (int i = tid; i < words; i += 128) d[i] = s[i];
  }
  {
    const int n = g.n_rows[u];
    const long long o = off[N1 + u];
    const float* s = g.formants + (size_t)row0 * 9;
    float* d = g.d_formants + (size_t)o * 9;
    for (int i = tid; i < n * 9; i += 128) d[i] = s[i];
    const float* s2 = g.energy + (size_t)row0 * 3;
    float* d2 = g.d_energy + (size_t)o * 3;
    for (int i = tid; i < n * 3; i += 128) d2[i] = s2[i];
  }
  {
    const int n = g.n_syls[u];
    const long long o = off[2 * N1 + u];
    const uint32_t* s = reinterpret_cast<const uint32_t*>(g.syls + sb);
    uint32_t* d = reinterpret_cast<uint32_t*>(g.d_syls + o);
    for (int i = tid; i < n * 4; i += 128) d[i] = s[i];
  }
  {
    const int n = g.n_feat[u];
    const long long o = off[3 * N1 + u];
    const double* s = g.features + (size_t)sb * FA_N_FEATURES;
    double* d = g.d_features + (size_t)o * FA_N_FEATURES;
    for (int i = tid; i < n * FA_N_FEATURES; i += 128) d[i] = s[i];
  }
}

}  // namespace

cudaError_t fa_launch_features(const FaFeatureParams& p, cudaStream_t s, int* launches) {
  if (p.utt_count <= 0) return cudaSuccess;
  fa_features_kernel<<<p.utt_count, kFeatThreads, 0, s>>>(p);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

cudaError_t fa_launch_prefix(const FaGatherArgs& a, cudaStream_t s, int* launches) {
  if (a.n_utt <= 0) return cudaSuccess;
  fa_prefix_kernel<<<1, 1024, 0, s>>>(a);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

cudaError_t fa_launch_gather(const FaGatherArgs& a, cudaStream_t s, int* launches) {
  if (a.n_utt <= 0) return cudaSuccess;
  fa_gather_kernel<<<a.n_utt, 128, 0, s>>>(a);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
