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
// frame, rows streamed with 16-byte loads straight into registers.  Integer compares only.
#include "fa_internal.cuh"

namespace {

constexpr int kPeakThreads = 128;

// NOTE (ptxas 12.9, sm_100a): with p.B / p.maxp read straight from the parameter bank, ptxas keeps B in a UNIFORM
// register (UR4) across the bin loop, but re-loads other parameters (maxp, B - 1) into the same UR4 inside the divergent
// emit() paths; the loop bound `B / 4` is then computed from whatever the last path left there (seen in the SASS:
// USHF.R.S32.HI UR4, URZ, 0x2, UR4 at the loop tail) -- rows lose their last iteration or the loop runs away.  Passing
// the two values through an opaque asm keeps them in ordinary per-thread registers, which sidesteps the problem.
// tests/test_gpu_parity.py checks g and the candidates of every frame.
// One thread per frame, no shared memory: the thread streams its own 4*B-byte row with 16-byte loads
// (rows are 512 B for B = 128: four full cache lines, every byte used; K1 has just written them, so most
// come from L2) and runs the automaton in registers.  200k frames = 6250 warps: one wave at full occupancy.
template <bool kStaged>
__global__ void __launch_bounds__(kPeakThreads, kStaged ? 10 : 12) fa_peaks_kernel(const FaPeaksParams p) {
  const long long fi = (long long)blockIdx.x * kPeakThreads + threadIdx.x;
  // (staged rows: the lanes of a warp load each other's rows, so lanes past the end stay until the scan is over)
  const bool live = fi < p.n_frames;
  if (!live && !(kStaged && (p.B & 31) == 0)) return;
  const long long f = p.row_begin + fi;
  int B = p.B, maxp = p.maxp;
  asm volatile("" : "+r"(B), "+r"(maxp));  // see the NOTE above
  const uint32_t* __restrict__ e = p.frames + (size_t)f * B;
  FaCand* out = p.cand + (size_t)f * maxp;
  int n = 0, lo = 0, pk = 0, hi = 0, flat = 0, dir = 0;
  unsigned long long g = 0;
  uint32_t epk = 0;  // e[pk]
  // exact prefix sums P[b] = e[0] + .. + e[b] at the raw bounds: pl = P[lo - 1], ph = P[hi] (K3 gets the
  // bandwidth energy of accumulate_fm @B35952 as ph - pl without touching the frame again)
  unsigned long long pre = 0, pl = 0, ph = 0;

  // Loop fission against divergence: the automaton runs thread-per-frame, so lanes close their peaks at different bins
  // and everything inside emit() is executed once per lane and peak.  emit() therefore only parks the raw candidate
  // (two 16-byte stores into its own output slot); the trim of lo / hi to bins >= e[pk]/10 -- data-dependent loops with
  // loads -- happens after the scan, candidate index by candidate index, all lanes of the warp together.
  auto emit = [&](int last) {
    if (n < maxp) {
      uint4* o4 = reinterpret_cast<uint4*>(out + n);  // two 16-byte stores into one sector
      o4[0] = make_uint4((uint32_t)lo | ((uint32_t)hi << 8) | ((uint32_t)pk << 16) | ((uint32_t)last << 24), epk,
                         (uint32_t)pl, (uint32_t)(pl >> 32));
      o4[1] = make_uint4((uint32_t)ph, (uint32_t)(ph >> 32), 0u, 0u);
    }
    n++;
  };
  auto step = [&](const int a, const uint32_t ea, const uint32_t e1, const uint32_t e2, const uint32_t e3) {
    g += ea;
    // here pre == P[a - 1]
    const bool rise = ea > e1 && (a < 2 || ea > e2) && (a < 3 || ea > e3);
    const bool fall = ea < e1 && (a < 2 || ea < e2) && (a < 3 || ea < e3);
    if (rise) {
      if (dir != 1) {
        if (dir == -1 && lo <= pk && pk < hi) emit(0);
        lo = a - 1;
        pl = pre - e1;
      }
      pk = a; epk = ea;
      dir = 1;
    } else if (fall) {
      if (dir != 0) { hi = a; ph = pre + ea; dir = -1; }
    } else if (dir == -1) {
      if (++flat > 2) {
        flat = 0;
        if (lo <= pk && pk < hi) emit(0);
        dir = 0;
      }
    } else if (dir == 1 && ea > e1) {
      pk = a; epk = ea;
    }
    if (a == B - 1 && dir == 1) {
      hi = a; pk = a; epk = ea;
      ph = pre + ea;
      if (lo < pk && pk <= hi) emit(1);
    }
    pre += ea;
  };

  uint32_t e1 = 0, e2 = 0, e3 = 0;  // e[a-1], e[a-2], e[a-3]
  if (kStaged && (B & 31) == 0) {
    // Rows through shared memory, 128 bytes (32 bins) of every row of the warp at a time: 8 lanes read one row chunk
    // contiguously (4 rows per load instruction, every sector requested once), then each lane scans its own row's chunk
    // from shared memory (row pitch 144 B: the eight 16-byte reads of a quarter-warp fall into disjoint banks).
    // ncu on the direct version: 63 % of the 30.9 M sectors requested from L2 were excess (a lane's 16-byte load opens a
    // 32-byte sector that is gone from L1 when the lane returns for the other half).  MEASURED (C2, B200): DRAM reads 322 ->
    // 223 MB, writes 128 -> 95 MB, but 0.245 ms against 0.158 ms for the direct version -- the scan is bound by its dependent
    // integer chain per bin and by occupancy (48 registers + 18 KB shared memory here), not by DRAM.  Kept as FA_K2_STAGED=1.
    __shared__ uint4 s_tile[kPeakThreads / 32][32][9];
    uint4 (*tile)[9] = s_tile[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, sub = lane & 7, rsel = lane >> 3;
    const long long warp_f0 = f - lane;                       // first frame of the warp
    const long long f_end = p.row_begin + p.n_frames;
    for (int c = 0; c < B / 32; c++) {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int row = 4 * k + rsel;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (warp_f0 + row < f_end)
          x = __ldg(reinterpret_cast<const uint4*>(p.frames + (size_t)(warp_f0 + row) * B + 32 * c) + sub);
        tile[row][sub] = x;
      }
      __syncwarp();
#pragma unroll 1
      for (int q = 0; q < 8; q++) {
        const uint4 x = tile[lane][q];
        const int a = 32 * c + 4 * q;
        if (a) step(a, x.x, e1, e2, e3); else pre = x.x;
        step(a + 1, x.y, x.x, e1, e2);
        step(a + 2, x.z, x.y, x.x, e1);
        step(a + 3, x.w, x.z, x.y, x.x);
        e3 = x.y; e2 = x.z; e1 = x.w;
      }
    }
  } else if ((B & 3) == 0) {
    const uint4* e4 = reinterpret_cast<const uint4*>(e);
    for (int q = 0; q < B / 4; q++) {
      const uint4 x = __ldg(e4 + q);
      const int a = 4 * q;
      if (q) step(a, x.x, e1, e2, e3); else pre = x.x;
      step(a + 1, x.y, x.x, e1, e2);
      step(a + 2, x.z, x.y, x.x, e1);
      step(a + 3, x.w, x.z, x.y, x.x);
      e3 = x.y; e2 = x.z; e1 = x.w;
    }
  } else {
    e1 = __ldg(e);
    pre = e1;
    for (int a = 1; a < B; a++) {
      const uint32_t ea = __ldg(e + a);
      step(a, ea, e1, e2, e3);
      e3 = e2; e2 = e1; e1 = ea;
    }
  }
  if (!live) return;
  p.ncand[f] = n;
  p.gsum[f] = (double)g;
  // trim (close() @B25717): while lo < pk and e[lo] < e[pk]/10: lo++; while hi > pk and e[hi] < e[pk]/10: hi--.
  // e[i] < e[pk]/10 in doubles  <=>  10*e[i] < e[pk] in integers (both exact).  The prefix sums follow the bounds.
  const int nc = n < maxp ? n : maxp;
  for (int c = 0; c < nc; c++) {
    uint4* o4 = reinterpret_cast<uint4*>(out + c);
    const uint4 a4 = o4[0];
    const uint2 b2 = *reinterpret_cast<const uint2*>(o4 + 1);
    int l2 = (int)(a4.x & 0xffu), h2 = (int)((a4.x >> 8) & 0xffu);
    const int pk2 = (int)((a4.x >> 16) & 0xffu);
    const unsigned long long top = a4.y;
    unsigned long long pl2 = a4.z | ((unsigned long long)a4.w << 32), ph2 = b2.x | ((unsigned long long)b2.y << 32);
    const int l0 = l2, h0 = h2;
    for (;;) {
      if (l2 >= pk2) break;
      const uint32_t x = __ldg(e + l2);
      if (!(10ull * x < top)) break;
      pl2 += x;
      l2++;
    }
    for (;;) {
      if (h2 <= pk2) break;
      const uint32_t x = __ldg(e + h2);
      if (!(10ull * x < top)) break;
      ph2 -= x;
      h2--;
    }
    if (l2 != l0 || h2 != h0) {
      o4[0] = make_uint4((a4.x & 0xffff0000u) | (uint32_t)l2 | ((uint32_t)h2 << 8), a4.y, (uint32_t)pl2, (uint32_t)(pl2 >> 32));
      *reinterpret_cast<uint2*>(o4 + 1) = make_uint2((uint32_t)ph2, (uint32_t)(ph2 >> 32));
    }
  }
}

}  // namespace

cudaError_t fa_launch_peaks(const FaPeaksParams& p, cudaStream_t s, int* launches) {
  if (p.n_frames <= 0) return cudaSuccess;
  const long long grid = (p.n_frames + kPeakThreads - 1) / kPeakThreads;
  if (p.staged) fa_peaks_kernel<true><<<(unsigned)grid, kPeakThreads, 0, s>>>(p);
  else fa_peaks_kernel<false><<<(unsigned)grid, kPeakThreads, 0, s>>>(p);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
