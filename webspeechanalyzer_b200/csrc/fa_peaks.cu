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


// ---------------------------------------------------------------------------------------------------------------------------
// K2 v2: the same automaton, thread per frame, but (a) the rows arrive through shared memory by 1-D TMA and (b) the step is
// branch-free.
//
// (a) A warp's 32 rows are 32 x 4B consecutive bytes of HBM.  Every lane issues ONE cp.async.bulk (UBLKCP) of its own row
//     into a padded shared-memory row (pitch 4B + 16 bytes: the eight 16-byte reads of a quarter-warp fall into disjoint
//     banks) and the warp waits on one mbarrier: every byte is fetched from HBM exactly once, asynchronously, with no register
//     staging -- the direct version above fetched each 32-byte sector twice (a lane's 16-byte load opens a sector that is gone
//     from L1 when it returns for the other half: 322 MB read for 102 MB of rows).
// (b) rise / fall make three-way data-dependent branches per bin, and with 32 independent frames per warp every path runs at
//     every bin (ncu: 112 warp-instructions per bin step).  Here the automaton's state (dir as two predicates, flat, lo, pk,
//     hi, the prefix sums) is updated by selects; only the 32-byte store of a closed peak stays behind a branch.
// Rows whose width is not a multiple of four bins cannot be bulk-copied (16-byte granularity): they are staged with plain
// coalesced loads into rows of odd pitch and scanned with 4-byte shared-memory reads (kVec = false).
constexpr int kPeak2Threads = 64;   // 2 warps: 33 KB (B = 128) / 65 KB (B = 256) of staged rows per CTA, 6 / 3 CTAs per SM

struct Scan {
  int n, lo, pk, hi, flat;
  bool d1, dm;                       // dir == 1, dir == -1
  uint32_t epk;
  unsigned long long pre, pl, ph;
};

template <bool kFirst>
__device__ __forceinline__ void scan_step(Scan& s, FaCand* out, const int maxp, const int a, const uint32_t ea, const uint32_t e1,
                                          const uint32_t e2, const uint32_t e3) {
  const bool up = ea > e1, dn = ea < e1;
  bool rise, fall;
  if (kFirst) {   // bins 1 and 2 compare with fewer neighbours (a < 2 / a < 3 in the reference loop)
    rise = up && (a < 2 || ea > e2) && (a < 3 || ea > e3);
    fall = dn && (a < 2 || ea < e2) && (a < 3 || ea < e3);
  } else {
    rise = up && ea > e2 && ea > e3;
    fall = dn && ea < e2 && ea < e3;
  }
  const bool neither = !rise && !fall;
  const int flat2 = s.flat + ((neither && s.dm) ? 1 : 0);
  const bool close = flat2 > 2;
  // a peak closes on the next rise after a fall, or after three flat bins; while dir == -1 the reference's guard
  // lo <= pk < hi always holds (lo, pk come from the rise that started the peak, hi from a later fall)
  if ((rise && s.dm) || close) {
    if (s.n < maxp) {
      uint4* o4 = reinterpret_cast<uint4*>(out + s.n);
      o4[0] = make_uint4((uint32_t)s.lo | ((uint32_t)s.hi << 8) | ((uint32_t)s.pk << 16), s.epk, (uint32_t)s.pl, (uint32_t)(s.pl >> 32));
      o4[1] = make_uint4((uint32_t)s.ph, (uint32_t)(s.ph >> 32), 0u, 0u);
    }
    s.n++;
  }
  const bool start = rise && !s.d1;
  const bool pkupd = rise || (neither && s.d1 && up);
  const bool fu = fall && (s.d1 || s.dm);
  const unsigned long long pre2 = s.pre + ea;
  s.lo = start ? a - 1 : s.lo;
  s.pl = start ? s.pre - e1 : s.pl;
  s.pk = pkupd ? a : s.pk;
  s.epk = pkupd ? ea : s.epk;
  s.hi = fu ? a : s.hi;
  s.ph = fu ? pre2 : s.ph;
  s.flat = close ? 0 : flat2;
  const bool d1n = rise || (s.d1 && !fall);
  s.dm = fu || (s.dm && !rise && !close);
  s.d1 = d1n;
  s.pre = pre2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool kVec>
__global__ void __launch_bounds__(kPeak2Threads) fa_peaks2_kernel(const FaPeaksParams p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int B = p.B, maxp = p.maxp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pitch = kVec ? 4 * B + 16 : 4 * (B | 1);                   // bytes per staged row
  unsigned char* tile = s_raw + (size_t)warp * 32 * pitch;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(s_raw + (size_t)(kPeak2Threads / 32) * 32 * pitch) + warp;
  const long long fi0 = ((long long)blockIdx.x * (kPeak2Threads / 32) + warp) * 32;   // first frame of the warp
  if (fi0 >= p.n_frames) return;
  const int rows = (int)(p.n_frames - fi0 < 32 ? p.n_frames - fi0 : 32);
  const long long f = p.row_begin + fi0 + lane;
  const bool live = lane < rows;
  const uint32_t* grow = p.frames + (size_t)f * B;
  if (kVec) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(rows * 4 * B) : "memory");
    }
    __syncwarp();
    if (live)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(tile + (size_t)lane * pitch)), "l"(grow), "r"(4 * B), "r"(smem_u32(bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\nselp.u32 %0, 1, 0, q;\n}"
                   : "=r"(done) : "r"(smem_u32(bar)) : "memory");
  } else {
    const uint32_t* g0 = p.frames + (size_t)(p.row_begin + fi0) * B;
    for (int i = lane; i < rows * B; i += 32) {
      const int r = i / B, c = i - r * B;
      reinterpret_cast<uint32_t*>(tile + (size_t)r * pitch)[c] = __ldg(g0 + i);
    }
    __syncwarp();
  }
  if (!live) return;
  const uint32_t* row = reinterpret_cast<const uint32_t*>(tile + (size_t)lane * pitch);
  FaCand* out = p.cand + (size_t)f * maxp;
  Scan s;
  s.n = 0; s.lo = 0; s.pk = 0; s.hi = 0; s.flat = 0; s.d1 = false; s.dm = false; s.epk = 0; s.pl = 0; s.ph = 0;
  uint32_t e1, e2, e3;
  const uint32_t e0 = row[0];
  s.pre = e0;
  if (kVec) {
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    {
      const uint4 x = r4[0];
      scan_step<true>(s, out, maxp, 1, x.y, x.x, 0u, 0u);
      scan_step<true>(s, out, maxp, 2, x.z, x.y, x.x, 0u);
      scan_step<false>(s, out, maxp, 3, x.w, x.z, x.y, x.x);
      e3 = x.y; e2 = x.z; e1 = x.w;
    }
#pragma unroll 2
    for (int q = 1; q < B / 4; q++) {
      const uint4 x = r4[q];
      const int a = 4 * q;
      scan_step<false>(s, out, maxp, a, x.x, e1, e2, e3);
      scan_step<false>(s, out, maxp, a + 1, x.y, x.x, e1, e2);
      scan_step<false>(s, out, maxp, a + 2, x.z, x.y, x.x, e1);
      scan_step<false>(s, out, maxp, a + 3, x.w, x.z, x.y, x.x);
      e3 = x.y; e2 = x.z; e1 = x.w;
    }
  } else {
    e1 = e0; e2 = 0; e3 = 0;
    for (int a = 1; a < B; a++) {
      const uint32_t ea = row[a];
      if (a < 3) scan_step<true>(s, out, maxp, a, ea, e1, e2, e3);
      else scan_step<false>(s, out, maxp, a, ea, e1, e2, e3);
      e3 = e2; e2 = e1; e1 = ea;
    }
  }
  // the frame ends while rising: the last bin closes the peak (lo < pk holds: lo <= B - 2)
  if (s.d1) {
    if (s.n < maxp) {
      uint4* o4 = reinterpret_cast<uint4*>(out + s.n);
      o4[0] = make_uint4((uint32_t)s.lo | ((uint32_t)(B - 1) << 8) | ((uint32_t)(B - 1) << 16) | (1u << 24), e1, (uint32_t)s.pl,
                         (uint32_t)(s.pl >> 32));
      o4[1] = make_uint4((uint32_t)s.pre, (uint32_t)(s.pre >> 32), 0u, 0u);
    }
    s.n++;
  }
  p.ncand[f] = s.n;
  p.gsum[f] = (double)(s.pre - e0);   // g = e[1] + .. + e[B-1], exact (< 2^53)
  // trim (close() @B25717): while lo < pk and e[lo] < e[pk]/10: lo++; while hi > pk and e[hi] < e[pk]/10: hi--.
  // e[i] < e[pk]/10 in doubles  <=>  10*e[i] < e[pk] in integers (both exact).  The prefix sums follow the bounds.  The row
  // is still in shared memory; the parked candidates come back from L2.
  // The (packed, amplitude) words of four candidates are fetched together (independent loads: one L2 round trip per four
  // candidates instead of one each -- ncu had 30 % of the kernel's stall samples on this read); a candidate that stands as
  // parked (3 of 4) costs nothing more.
  const int nc = s.n < maxp ? s.n : maxp;
  for (int c0 = 0; c0 < nc; c0 += 4) {
    uint2 hd[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (c0 + k < nc) hd[k] = *reinterpret_cast<const uint2*>(out + c0 + k);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (c0 + k >= nc) break;
      int l2 = (int)(hd[k].x & 0xffu), h2 = (int)((hd[k].x >> 8) & 0xffu);
      const int pk2 = (int)((hd[k].x >> 16) & 0xffu);
      const uint32_t top = hd[k].y;
      uint32_t thr = top / 10u;          // 10 x < top  <=>  x < ceil(top / 10)
      thr += thr * 10u != top;
      const bool tl = l2 < pk2 && row[l2] < thr, th = h2 > pk2 && row[h2] < thr;
      if (!tl && !th) continue;
      uint4* o4 = reinterpret_cast<uint4*>(out + c0 + k);
      const uint2 a2 = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(o4) + 8);
      const uint2 b2 = *reinterpret_cast<const uint2*>(o4 + 1);
      unsigned long long pl2 = a2.x | ((unsigned long long)a2.y << 32), ph2 = b2.x | ((unsigned long long)b2.y << 32);
      for (;;) {
        if (l2 >= pk2) break;
        const uint32_t x = row[l2];
        if (!(x < thr)) break;
        pl2 += x;
        l2++;
      }
      for (;;) {
        if (h2 <= pk2) break;
        const uint32_t x = row[h2];
        if (!(x < thr)) break;
        ph2 -= x;
        h2--;
      }
      o4[0] = make_uint4((hd[k].x & 0xffff0000u) | (uint32_t)l2 | ((uint32_t)h2 << 8), top, (uint32_t)pl2, (uint32_t)(pl2 >> 32));
      *reinterpret_cast<uint2*>(o4 + 1) = make_uint2((uint32_t)ph2, (uint32_t)(ph2 >> 32));
    }
  }
}

}  // namespace

cudaError_t fa_launch_peaks(const FaPeaksParams& p, cudaStream_t s, int* launches) {
  if (p.n_frames <= 0) return cudaSuccess;
  if (p.staged >= 0) {   // FA_K2_IMPL: -1 = the direct version (A/B), 0 = v2
    const bool vec = (p.B & 3) == 0;
    const int pitch = vec ? 4 * p.B + 16 : 4 * (p.B | 1);
    const size_t smem = (size_t)(kPeak2Threads / 32) * 32 * pitch + 8 * (kPeak2Threads / 32);
    const long long warps = (p.n_frames + 31) / 32;
    const long long grid = (warps + kPeak2Threads / 32 - 1) / (kPeak2Threads / 32);
    static bool attr_done[2] = {false, false};
    cudaError_t e = cudaSuccess;
    if (vec) {
      if (!attr_done[0]) { e = cudaFuncSetAttribute(fa_peaks2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done[0] = true; }
      if (e != cudaSuccess) return e;
      fa_peaks2_kernel<true><<<(unsigned)grid, kPeak2Threads, smem, s>>>(p);
    } else {
      if (!attr_done[1]) { e = cudaFuncSetAttribute(fa_peaks2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done[1] = true; }
      if (e != cudaSuccess) return e;
      fa_peaks2_kernel<false><<<(unsigned)grid, kPeak2Threads, smem, s>>>(p);
    }
    if (launches) (*launches)++;
    return cudaGetLastError();
  }
  const long long grid = (p.n_frames + kPeakThreads - 1) / kPeakThreads;
  fa_peaks_kernel<false><<<(unsigned)grid, kPeakThreads, 0, s>>>(p);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
