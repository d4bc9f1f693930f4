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
// frame; rows staged in shared memory by 1-D TMA, a branch-free step.  Integer compares only.
#include "fa_internal.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------------------
// The automaton runs thread per frame; (a) the rows arrive through shared memory by 1-D TMA and (b) the step is branch-free.
// (Round 1 streamed every lane's row with 16-byte global loads and branched three ways per bin: 0.157 ms on C2; this: 0.122.)
//
// (a) A warp's 32 rows are 32 x 4B consecutive bytes of HBM.  Every lane issues ONE cp.async.bulk (UBLKCP) of its own row
//     into a padded shared-memory row (pitch 4B + 16 bytes: the eight 16-byte reads of a quarter-warp fall into disjoint
//     banks) and the warp waits on one mbarrier: every byte is fetched from HBM exactly once, asynchronously, with no register
//     staging -- the round-1 version fetched each 32-byte sector twice (a lane's 16-byte load opens a sector that is gone
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
