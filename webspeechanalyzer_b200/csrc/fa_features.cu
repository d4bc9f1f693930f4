// fa_features.cu -- K4: the 53 statistics per segment (level 5) or per syllable (level 13), and
// K5: gather of the per-utterance tables into dense arrays.
//
// Restates formant_features (/root/reference/dist/main.js:2@B32369, vector assembly @B33436) and the
// stats helpers it calls (/root/reference/src/stats.js:29-64: array_mean_NZ, only_std_NZ,
// mean_std_NZ, arraySum).  A row is a segmented reduction over its frames: one warp per row; the
// 20*log10(E) of every (frame, slot) is computed lane-parallel, then one lane per formant slot walks
// the frames in array order (the reference's summation order, so the doubles come out bit-identical
// to the oracle's) in two passes -- sums and the accent automaton first, squared deviations second.
// FP64 throughout, log10 from include/fa_jsmath.h.  Kilobytes per row: latency bound, not HBM bound.
#include "fa_internal.cuh"
#include "fa_jsmath.h"

namespace {

constexpr int kFeatThreads = 64;

constexpr int kChunk = 128;   // frames per chunk: dB values of 3 slots x 128 frames + the staged rows in shared memory per warp (7.7 KB:
                              // with 256 the 31 KB per CTA capped the resident CTAs and a 36 k-row shard ran 15 % slower)

struct SlotAcc {
  double cnt, runs, up, down, sum_c, sum_w, sum_T, sum_k, sum_knz, sum_M, sum_Anz, L;
  double acc_w, acc_k, acc_a, mean_w, dbm, accm;
  int m, n_knz, na, n_anz, S;
  bool prev;
};

// Per row: the expensive part (20*log10(E), fdlibm, ~150 dependent FP64 ops) is done lane-per-frame into
// shared memory; the order-sensitive part (sums in array order, run / jump / accent automaton) is done by three
// lanes, one per formant slot, over the precomputed values.  Two passes: sums, then squared deviations.
// The walkers read the chunk's rows from SHARED memory (lane-parallel, coalesced staging of the 9-float rows next to the dB
// values) instead of loading F[row] from global memory inside the loop: 0.109 -> 0.098 ms on C2, 0.45 -> 0.41 ms on 24 k syllable rows.  Measured and rejected in round
// 2: ten rows per warp through a flattened work list (lanes 3g .. 3g + 2 walk row g: 0.37 ms -- the ten rows' log10 phases
// queue up in one warp) and one row per CTA with each walk in its own warp (0.116 ms): the walk is a chain of dependent FP64
// compares and branches, ~500 cycles per frame and pass, whatever the mapping.
__device__ void row_features(const float* __restrict__ F, const int len, const double ymax, double (*sdb)[kChunk],
                             float* __restrict__ sF /* [(kChunk + 1) * 9]: row -1 of the chunk first */,
                             const int lane, double* __restrict__ out /* 48 = 3 x 16 */) {
  SlotAcc A;
  A.cnt = A.runs = A.up = A.down = A.sum_c = A.sum_w = A.sum_T = A.sum_k = A.sum_knz = A.sum_M = A.sum_Anz = A.L = 0;
  A.acc_w = A.acc_k = A.acc_a = A.mean_w = A.dbm = A.accm = 0;
  A.m = A.n_knz = A.na = A.n_anz = A.S = 0;
  A.prev = false;
  const int n = lane;  // slot handled by this lane in the sequential phases (lanes 0..2)
  for (int pass = 0; pass < 2; pass++) {
    bool prev = false;
    int S = 0;
    double L = 0;
    for (int c0 = 0; c0 < len; c0 += kChunk) {
      const int cn = min(kChunk, len - c0);
      __syncwarp();
      // rows c0 - 1 .. c0 + cn - 1 of the segment (the row in front of the chunk feeds the first jump test)
      for (int i = lane + (c0 > 0 ? 0 : 9); i < 9 * (cn + 1); i += 32) sF[i] = F[(size_t)c0 * 9 - 9 + i];
      __syncwarp();
      for (int i = lane; i < 3 * cn; i += 32) {
        const int sl = i / cn, t = i - sl * cn;
        const double a = (double)sF[(size_t)(t + 1) * 9 + 3 * sl + 1];
        sdb[sl][t] = a > 0 ? 20.0 * fa_js_log10(a) : 0.0;
      }
      __syncwarp();
      if (lane < 3) {
        for (int t = 0; t < cn; t++) {
          const int row = (t + 1) * 9 + 3 * n;
          const double r = (double)sF[row], a = (double)sF[row + 1];
          if (r > 0 && a > 0) {
            const double d = sdb[n][t];
            if (pass == 0) {
              const double f = (double)sF[row + 2];
              A.sum_c += r * d; A.sum_w += r; A.sum_M += f * d; A.sum_T += a; A.sum_k += d;
              if (d > 0) { A.sum_knz += d; A.n_knz++; }
              A.m++;
            } else {
              const double dw = r - A.mean_w, dk = d - A.dbm;
              A.acc_w += dw * dw;
              A.acc_k += dk * dk;
            }
            if (prev) {
              if (pass == 0) {
                const double j = r - (double)sF[row - 9];
                if (j > 1) A.up += j; else if (j < -1) A.down += -1 * j;
              }
              if (a > L) { L = a; S = 1; }
              else if (S == 1 && a < L / 2) {
                if (L > 10) {
                  if (pass == 0) { A.na++; if (d > 0) { A.sum_Anz += d; A.n_anz++; } }
                  else { const double da = d - A.accm; A.acc_a += da * da; }
                }
                L = 0; S = -1;
              }
            }
            if (pass == 0) {
              if (!prev) A.runs += 1;
              A.cnt += 1;
            }
            prev = true;
          } else { prev = false; S = 0; L = 0; }
        }
      }
    }
    if (pass == 0 && lane < 3 && A.runs > 0) {
      A.mean_w = A.sum_w / (double)A.m;  // every bin in the list is > 0
      A.dbm = A.sum_knz / (double)A.n_knz;
      if (A.na > 0) A.accm = A.sum_Anz / (double)A.n_anz;
    }
  }
  if (lane < 3) {
    double fmean = 0, fstd = 0, dbm = 0, dbs = 0, e_len = 0, e_cnt = 0, span = 0, accm = 0, accs = 0, prom = 0;
    if (A.runs > 0) {
      e_len = A.sum_T / (double)len * 100 / ymax;
      e_cnt = A.sum_T / A.cnt * 100 / ymax;
      fmean = A.sum_c / A.sum_k;
      span = A.sum_M / A.sum_k;
      dbm = A.dbm;
      fstd = fa_sqrt(A.acc_w / (double)A.m);
      dbs = fa_sqrt(A.acc_k / (double)A.m);
      if (A.na > 0) {
        accm = A.accm;
        accs = fa_sqrt(A.acc_a / (double)A.na);
        prom = 100 * (accm / (A.sum_k / (double)A.m) - 1);
      }
    }
    double* o = out + 16 * n;
    o[0] = fmean; o[1] = fstd; o[2] = dbm; o[3] = dbs; o[4] = e_len; o[5] = e_cnt; o[6] = span;
    o[7] = A.cnt; o[8] = A.runs; o[9] = A.up; o[10] = A.down; o[11] = (double)A.na; o[12] = accm; o[13] = accs;
    o[14] = prom; o[15] = 100 * A.cnt / (double)len;
  }
}

__global__ void __launch_bounds__(kFeatThreads) fa_features_kernel(const FaFeatureParams p) {
  __shared__ double s_db[kFeatThreads / 32][3][kChunk];
  __shared__ float s_F[kFeatThreads / 32][(kChunk + 1) * 9];
  const int u = p.utt_begin + blockIdx.x;
  const long long row0 = p.frame_off[u], sb = row0 + u;
  const int nseg = p.n_segs[u];
  const bool per_syl = p.level == 13;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // level 5: rows are the stored segments, in seg_ci order
  __shared__ int s_rows;
  if (threadIdx.x == 0) {
    int c = 0;
    if (per_syl) c = p.n_syls[u];
    else for (int s = 0; s < nseg; s++) c += p.segs[sb + s].stored >= 0;
    s_rows = c;
    p.n_feat[u] = c;
  }
  __syncthreads();
  const int R = s_rows;
  for (int row = warp + (int)blockIdx.y * (kFeatThreads / 32); row < R; row += (kFeatThreads / 32) * (int)gridDim.y) {
    const float* F;
    int len;
    const fa_segment* sg;
    if (per_syl) {
      const fa_syllable sy = p.syls[sb + row];
      int s = 0;
      while (p.segs[sb + s].stored != sy.stored_seg) s++;  // stored indices increase in seg_ci order
      sg = &p.segs[sb + s];
      F = p.formants + (size_t)(row0 + (p.epochs ? p.epochs[sb + s].first : sg->row_offset) + sy.start) * 9;
      len = sy.len;
    } else {
      int s = 0, c = -1;
      for (;; s++) { if (p.segs[sb + s].stored >= 0 && ++c == row) break; }
      sg = &p.segs[sb + s];
      F = p.formants + (size_t)(row0 + (p.epochs ? p.epochs[sb + s].first : sg->row_offset)) * 9;
      len = sg->len;
    }
    double* out = p.features + (size_t)(sb + row) * FA_N_FEATURES;
    if (lane == 3) {
      out[0] = (double)len;
      out[1] = fa_sqrt((double)len);
      out[2] = sg->cs_ratio;
      out[3] = fa_js_log10(sg->ymax);
      out[4] = sg->vmin;
    }
    row_features(F, len, sg->ymax, s_db[warp], s_F[warp], lane, out + 5);
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
  if (g.l3_mult) {   // level 3: fa_track_point rows (6 words) from the point-pool base; no energy rows
    const int n = g.n_rows[u];
    const long long o = off[N1 + u];
    const uint32_t* s = reinterpret_cast<const uint32_t*>(g.formants) + (size_t)row0 * g.l3_mult * 6;
    uint32_t* d = reinterpret_cast<uint32_t*>(g.d_formants) + (size_t)o * 6;
    for (int i = tid; i < n * 6; i += 128) d[i] = s[i];
  } else if (g.epochs) {   // stream mode: segment by segment from the epochs' frame ranges, segments dealt over the row slices
    const long long o = off[N1 + u];
    const int nseg = g.n_segs[u];
    for (int sgi = blockIdx.y; sgi < nseg; sgi += gridDim.y) {
      const fa_segment sg = g.segs[sb + sgi];
      if (sg.stored < 0) continue;
      const size_t src = (size_t)(row0 + g.epochs[sb + sgi].first), dst = (size_t)(o + sg.row_offset);
      const float* s = g.formants + src * 9;
      float* d = g.d_formants + dst * 9;
      for (int i = tid; i < sg.len * 9; i += 128) d[i] = s[i];
      const float* s2 = g.energy + src * 3;
      float* d2 = g.d_energy + dst * 3;
      for (int i = tid; i < sg.len * 3; i += 128) d2[i] = s2[i];
    }
    if (blockIdx.y != 0) return;
  } else {
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
    const uint32_t* s = reinterpret_cast<const uint32_t*>(g.syls + (g.l3_mult ? row0 * g.l3_mult : sb));
    uint32_t* d = reinterpret_cast<uint32_t*>(g.d_syls + o);
    for (int i = tid; i < n * 4; i += 128) d[i] = s[i];
  }
  {
    const int n = g.n_feat[u];
    const long long o = off[3 * N1 + u];
    const int W = g.feat_width;
    const double* s = g.features + (size_t)(g.feat_base ? g.feat_base[u] : sb) * W;
    double* d = g.d_features + (size_t)o * W;
    for (int i = tid; i < n * W; i += 128) d[i] = s[i];
  }
}

}  // namespace

cudaError_t fa_launch_features(const FaFeatureParams& p, cudaStream_t s, int* launches) {
  if (p.utt_count <= 0) return cudaSuccess;
  fa_features_kernel<<<dim3(p.utt_count, p.row_slices > 0 ? p.row_slices : 1), kFeatThreads, 0, s>>>(p);
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
  fa_gather_kernel<<<dim3(a.n_utt, a.epochs && a.row_slices > 1 ? a.row_slices : 1), 128, 0, s>>>(a);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
