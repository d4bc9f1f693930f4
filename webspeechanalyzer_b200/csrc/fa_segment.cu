// fa_segment.cu -- K3: the per-utterance sequential scan (stages S2b, S3, S3b, S3c).
//
// Restates, for one utterance per warp (utterances are independent: reset_segmentation @B25053),
//   D()  voiced/pause decision + segment state machine   /root/reference/dist/main.js:2@B26571, @B26663
//   C()  adaptive noise gate                              @B28506
//   L()  segment reset                                    @B25649
//   O()  segment finalisation                             @B27088
//   accumulate_fm + score                                 @B35952, @B37340
//   get_ranked_formants / straighten_formants             @B35670, @B35074
//   sep_syllables                                         @B34757
// The state machine is strictly sequential in time (the threshold v of frame t depends on frame
// t-1), so the time loop is serial and the 32 lanes parallelise what is parallel inside a frame:
// candidate filtering, track x peak scoring (lane per live track, warp arg-max per peak), the
// bandwidth energy sums, and at finalisation the ranking (count-smaller sort), the per-row
// application of track points (lane per frame row) and the syllable bit scan.
// All decisions are integer compares or IEEE double arithmetic without contraction
// (--fmad=false) and the V8-equivalent log10/pow of include/fa_jsmath.h, so boundaries are
// bit-exact against oracle/fa_oracle.c.
//
// This kernel moves kilobytes per utterance; it is latency bound, not HBM bound.
#include <cstdlib>

#include "fa_internal.cuh"
#include "fa_jsmath.h"

namespace {

constexpr int kWarps = 2;   // 47 KB of shared memory per CTA: fits beside two K1a CTAs, so sub-batches overlap
constexpr int ACAP = 128;  // live-track slots per utterance (tracks with lastFrame >= c_ci - 3)
constexpr int PCAP = 136;  // accepted peaks per frame (>= maxp = B/2 + 4)
constexpr unsigned FULL = 0xffffffffu;
constexpr int BIG = 0x7fffffff;

constexpr int CMAX = 10;   // candidate peaks per live track: a +-8 bin window holds <= 9 peaks (>= 2 bins apart)

struct WarpShared {
  uint32_t e[FA_MAX_BANDS];
  unsigned long long P[FA_MAX_BANDS];      // inclusive prefix sums of e (exact)
  uint32_t pmask[FA_MAX_BANDS / 32 + 1];   // bit b: an accepted peak has pk == b
  unsigned char pidx[FA_MAX_BANDS];        // its index in the accepted list
  unsigned char plo[PCAP], phi[PCAP], ppk[PCAP];
  int owner[PCAP];
  unsigned long long bestbits[PCAP];       // best score per peak (bit pattern of a positive double)
  int a_id[ACAP], a_last_frame[ACAP], a_last_bin[ACAP], a_npts[ACAP], a_b2[ACAP], a_b3[ACAP];
  uint32_t a_last_amp[ACAP];
  double a_vel[ACAP], a_sum_e[ACAP], a_sum_eb[ACAP];
  double cand_sc[CMAX][ACAP];              // [k][slot]: lane-consecutive slots -> conflict free
  unsigned char cand_o[CMAX][ACAP];
  unsigned char cand_n[ACAP];              // candidates of the slot this frame
};

struct ScanState {
  int current_frame, no_fm_segs, c_ci, c_started, w, k;
  double y, v, x, v0, T, s_energy, c_energy;
  int n_tr, n_act, n_pts;
  int n_segs, n_stored, n_rows, n_syls;
  int overflow;
};

struct Bases {
  long long row0;   // first frame row of the utterance
  long long tb;     // track table base
  long long pb;     // point pool base
  long long sb;     // segment / syllable / row-scratch base (row0 + u)
  int F, tcap;
};

__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// _() @B37340
__device__ __forceinline__ double fm_score(int gap, double dist, int count, int bin_old, int bin_new, double amp_old,
                                           double amp_new, double velocity) {
  double s;
  if (amp_old >= amp_new) s = amp_new / amp_old;
  else {
    if (!(amp_new > 0)) return 0;
    s = amp_old / amp_new;
  }
  if (gap == 0) return s > 0.1 ? 300.0 * s / dist : 0;
  if (s < 0.001) return 0;
  if (s >= 1) s = 10; else if (s < 0.1) s = 1; else s *= 10;
  double t = 10.0 - fabs((double)bin_new - (double)bin_old - velocity);
  if (t < 0) return 0;
  if (t < 1) t = 1;
  const int i = count > 10 ? 10 : count;
  return 10.0 / (double)gap * (t * t + (double)i * s);
}

// L() @B25649 (+ clear_fm @B35919)
__device__ __forceinline__ void seg_reset(ScanState& st, int started) {
  st.c_ci = 0;
  st.c_started = started;
  st.no_fm_segs = 0;
  st.n_tr = 0;
  st.n_act = 0;
  st.n_pts = 0;
  st.s_energy = 0.0;
  st.c_energy = 0.0;
}

// C() @B28506
__device__ __noinline__ void noise_gate(ScanState& st, double e) {
  st.w++;
  if (e > st.y || (st.w > 40 && e > 2 * st.v)) {
    if (e >= st.y) { st.w = 0; st.x = st.y = e; }
    else if (e > st.x / 100) { st.y -= fa_js_parse_int(st.y / 8); st.w = 35; }
    const double t = fa_js_log10(st.y);
    if (t > 7) st.v = fa_js_parse_int(fa_js_pow(10, t - 3) / 20);
    else if (t > 6) st.v = fa_js_parse_int(fa_js_pow(10, t - 3) / 2);
    else if (t > 4) st.v = fa_js_parse_int(fa_js_pow(10, t - 2) / 2);
    else if (t > 2) st.v = fa_js_parse_int(fa_js_pow(10, t / 3));
    else if (t > 1) st.v = fa_js_parse_int(st.y / 10);
    else st.v = 1;
    st.v0 = st.v;
    if (st.k > 0 && st.T / (double)st.k < 30 * st.v) { seg_reset(st, 0); st.k = 0; st.T = 0; }
    st.T += st.y;
    st.k += 1;
  } else if (st.v > 10 && st.v > st.v0 / 10 && st.w > 20) {
    st.v -= fa_js_parse_int(st.v0 / 20);
    if (st.v < 10) st.v = 10;
  }
}

// accumulate_fm @B35952
__device__ __noinline__ void accumulate_fm(const FaSegmentParams& p, WarpShared& S, ScanState& st, const Bases& bs,
                                           const int n_peaks, const int n_label, const double g, const double vmin,
                                           const int lane) {
  if (n_peaks < 1) return;
  st.s_energy += g;
  // drop tracks that can no longer match (gap >= 4); the reference keeps them but never touches them again
  {
    int n_new = 0;
    for (int r0 = 0; r0 < st.n_act; r0 += 32) {
      const int r = r0 + lane;
      const bool valid = r < st.n_act;
      int id = 0, lf = 0, lb = 0, np = 0, b2 = 0, b3 = 0;
      uint32_t la = 0;
      double vel = 0, se = 0, seb = 0;
      if (valid) {
        id = S.a_id[r]; lf = S.a_last_frame[r]; lb = S.a_last_bin[r]; np = S.a_npts[r]; b2 = S.a_b2[r];
        b3 = S.a_b3[r]; la = S.a_last_amp[r]; vel = S.a_vel[r]; se = S.a_sum_e[r]; seb = S.a_sum_eb[r];
      }
      const bool keep = valid && (n_label - lf < 4);
      if (valid && !keep) {
        p.trk_count[bs.tb + id] = np;
        p.trk_sum_e[bs.tb + id] = se;
        p.trk_sum_eb[bs.tb + id] = seb;
      }
      const unsigned m = __ballot_sync(FULL, keep);
      __syncwarp();
      if (keep) {
        const int d = n_new + __popc(m & ((1u << lane) - 1));
        S.a_id[d] = id; S.a_last_frame[d] = lf; S.a_last_bin[d] = lb; S.a_npts[d] = np; S.a_b2[d] = b2;
        S.a_b3[d] = b3; S.a_last_amp[d] = la; S.a_vel[d] = vel; S.a_sum_e[d] = se; S.a_sum_eb[d] = seb;
      }
      n_new += __popc(m);
      __syncwarp();
    }
    st.n_act = n_new;
  }
  // per-frame tables: peak bitmask + index, exact prefix sums of the frame
  const int B = p.B;
  {
    const int nw = (B + 31) >> 5;
    if (lane <= nw) S.pmask[lane] = 0u;
    const int C = (B + 31) >> 5;            // bins per lane
    const int b0 = lane * C, b1 = min(b0 + C, B);
    unsigned long long loc = 0;
    for (int b = b0; b < b1; b++) loc += S.e[b];
    unsigned long long incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    unsigned long long run = incl - loc;
    for (int b = b0; b < b1; b++) { run += S.e[b]; S.P[b] = run; }
    __syncwarp();
    for (int o = lane; o < n_peaks; o += 32) {
      const int pk = S.ppk[o];
      atomicOr(&S.pmask[pk >> 5], 1u << (pk & 31));
      S.pidx[pk] = (unsigned char)o;
      S.owner[o] = BIG;
      S.bestbits[o] = 0ull;
    }
    __syncwarp();
  }
  // (1) score every (live track, peak within its window): lane per track.  best[o] = max score via
  //     shared-memory atomicMax on the bit pattern; ties go to the earlier track via atomicMin on the slot.
  // (live tracks: 17 on average, rarely more than 40 -> usually one pass of 32 lanes)
  const int n_pass = (st.n_act + 31) >> 5;
  bool ovf = false;
  for (int ps = 0; ps < n_pass; ps++) {
    const int r = ps * 32 + lane;
    if (r < st.n_act) {
      int cnt = 0;
      const int gap = n_label - S.a_last_frame[r];
      if (gap >= 0 && gap < 4) {
        const int lb = S.a_last_bin[r];
        const int lim = gap == 0 ? 3 : gap == 1 ? 4 : gap == 2 ? 6 : 9;  // DIST @B32325
        const int wlo = max(lb - lim + 1, 0), whi = min(lb + lim - 1, B - 1);
        const int word = wlo >> 5, sh = wlo & 31;
        const unsigned long long two = (unsigned long long)S.pmask[word] | ((unsigned long long)S.pmask[word + 1] << 32);
        unsigned bits = (unsigned)(two >> sh) & ((2u << (whi - wlo)) - 1u);
        const double amp_old = (double)S.a_last_amp[r], vel = S.a_vel[r];
        const int np = S.a_npts[r];
        while (bits) {
          const int j = __ffs(bits) - 1;
          bits &= bits - 1;
          const int bin = wlo + j;
          const double sc = fm_score(gap, (double)abs(lb - bin), np, lb, bin, amp_old, (double)S.e[bin], vel);
          if (sc > 1) {
            const int o = S.pidx[bin];
            if (cnt < CMAX) {
              S.cand_sc[cnt][r] = sc;
              S.cand_o[cnt][r] = (unsigned char)o;
            }
            cnt++;
            atomicMax(&S.bestbits[o], (unsigned long long)__double_as_longlong(sc));
          }
        }
      }
      ovf |= cnt > CMAX;
      S.cand_n[r] = (unsigned char)(cnt > CMAX ? CMAX : cnt);
    }
  }
  if (__any_sync(FULL, ovf)) { st.overflow = 1; return; }
  __syncwarp();
  for (int ps = 0; ps < n_pass; ps++) {
    const int r = ps * 32 + lane;
    const int cnt = r < st.n_act ? S.cand_n[r] : 0;
    for (int k = 0; k < cnt; k++) {
      const int o = S.cand_o[k][r];
      if ((unsigned long long)__double_as_longlong(S.cand_sc[k][r]) == S.bestbits[o]) atomicMin(&S.owner[o], r);
    }
  }
  __syncwarp();
  // (2) every owning track absorbs its (merged) peaks: lane per track, its owned peaks are among its candidates
  unsigned long long moved = 0;
  for (int ps = 0; ps < n_pass; ps++) {
    const int r = ps * 32 + lane;
    int first = -1, lo = 0, hi = 0, o_bin = 0;
    uint32_t bamp = 0;
    const int my_cnt = r < st.n_act ? S.cand_n[r] : 0;
    for (int k = 0; k < my_cnt; k++) {
      const int o = S.cand_o[k][r];
      if (S.owner[o] == r) {
        const int pk = S.ppk[o];
        const uint32_t a = S.e[pk];
        if (first < 0) { first = o; lo = S.plo[o]; hi = S.phi[o]; o_bin = pk; bamp = a; }
        else {
          lo = min(lo, (int)S.plo[o]);
          hi = max(hi, (int)S.phi[o]);
          if (a > bamp) { bamp = a; o_bin = pk; }
        }
      }
    }
    bool upd = false;
    uint32_t amp0 = 0;
    if (first >= 0) {
      amp0 = S.e[S.ppk[first]];
      upd = (double)amp0 > vmin;
    }
    const unsigned um = __ballot_sync(FULL, upd);
    if (upd) {
      const unsigned long long Ei = S.P[hi] - (lo > 0 ? S.P[lo - 1] : 0ull);
      const double E = (double)Ei;
      moved += Ei;
      const int h = S.a_npts[r];
      const int b1 = S.a_last_bin[r], b2 = S.a_b2[r], b3 = S.a_b3[r];
      double vel = S.a_vel[r];
      if (h >= 3) vel = (double)(o_bin - b1 + (b2 - b1) + (b3 - b2)) / 3;
      else if (h == 2) vel = (double)(o_bin - b1 + (b2 - b1)) / 2;
      else if (h == 1) vel = (double)(o_bin - b1);
      S.a_vel[r] = vel; S.a_last_frame[r] = n_label; S.a_b3[r] = b2; S.a_b2[r] = b1; S.a_last_bin[r] = o_bin;
      S.a_last_amp[r] = amp0; S.a_npts[r] = h + 1; S.a_sum_e[r] += E; S.a_sum_eb[r] += E * (double)o_bin;
      const long long q = bs.pb + st.n_pts + __popc(um & ((1u << lane) - 1));
      p.pt_track[q] = S.a_id[r]; p.pt_ord[q] = h; p.pt_frame[q] = n_label;
      p.pt_binspan[q] = o_bin | ((hi - lo + 1) << 16); p.pt_e[q] = E;
    }
    st.n_pts += __popc(um);
  }
  {
    const double mv = (double)warp_sum_u64(moved);  // exact integers: one subtraction == the reference's sequence
    st.s_energy -= mv;
    st.c_energy += mv;
  }
  __syncwarp();
  // (3) un-owned peaks above the gate start new tracks, in peak order
  for (int o0 = 0; o0 < n_peaks; o0 += 32) {
    const int o = o0 + lane;
    const bool valid = o < n_peaks;
    const int pk = valid ? S.ppk[o] : 0;
    const uint32_t amp = valid ? S.e[pk] : 0;
    const bool mk = valid && S.owner[o] == BIG && (double)amp > vmin;
    const unsigned m = __ballot_sync(FULL, mk);
    const int cnt = __popc(m);
    if (st.n_act + cnt > ACAP || st.n_tr + cnt > bs.tcap) { st.overflow = 1; return; }
    if (mk) {
      const int pos = __popc(m & ((1u << lane) - 1));
      const int slot = st.n_act + pos, id = st.n_tr + pos;
      const int lo = S.plo[o], hi = S.phi[o];
      const double E = (double)(S.P[hi] - (lo > 0 ? S.P[lo - 1] : 0ull));
      S.a_id[slot] = id; S.a_last_frame[slot] = n_label; S.a_last_bin[slot] = pk; S.a_last_amp[slot] = amp;
      S.a_vel[slot] = 0; S.a_npts[slot] = 1; S.a_b2[slot] = 0; S.a_b3[slot] = 0; S.a_sum_e[slot] = E;
      S.a_sum_eb[slot] = E * (double)pk;
      const long long q = bs.pb + st.n_pts + pos;
      p.pt_track[q] = id; p.pt_ord[q] = 0; p.pt_frame[q] = n_label; p.pt_binspan[q] = pk | ((hi - lo + 1) << 16);
      p.pt_e[q] = E;
    }
    st.n_act += cnt;
    st.n_tr += cnt;
    st.n_pts += cnt;
  }
  __syncwarp();
}

// O() @B27088.  Returns 1 stored, 0 ignored, -1 rejected (where the JS throws inside straighten_formants).
__device__ __noinline__ int finalize_segment(const FaSegmentParams& p, WarpShared& S, ScanState& st, const Bases& bs,
                                             const int n_arg, const int lane) {
  const int len = n_arg - st.no_fm_segs;
  if (!(len > p.seg_min_frames && st.c_started >= 2)) return 0;
  const int start = st.current_frame - len;
  const double vmin = st.v;
  // flush live tracks
  for (int r = lane; r < st.n_act; r += 32) {
    const int id = S.a_id[r];
    p.trk_count[bs.tb + id] = S.a_npts[r];
    p.trk_sum_e[bs.tb + id] = S.a_sum_e[r];
    p.trk_sum_eb[bs.tb + id] = S.a_sum_eb[r];
  }
  __syncwarp();
  const int T = st.n_tr;
  // get_ranked_formants @B35670: count >= 2, mean >= 7, stable ascending by mean
  int nr = 0;
  for (int i0 = 0; i0 < T; i0 += 32) {
    const int i = i0 + lane;
    bool elig = false;
    if (i < T) {
      const double m = p.trk_sum_eb[bs.tb + i] / p.trk_sum_e[bs.tb + i];
      elig = p.trk_count[bs.tb + i] >= 2 && m >= 7;
      p.trk_mean[bs.tb + i] = elig ? m : -1.0;
      p.trk_slot[bs.tb + i] = -1;
    }
    nr += __popc(__ballot_sync(FULL, elig));
  }
  __syncwarp();
  for (int i = lane; i < T; i += 32) {
    const double m = p.trk_mean[bs.tb + i];
    if (m >= 7) {
      int rank = 0;
      for (int j = 0; j < T; j++) {
        const double mj = p.trk_mean[bs.tb + j];
        if (mj >= 7 && (mj < m || (mj == m && j < i))) rank++;
      }
      p.trk_rank[bs.tb + i] = rank;
      p.trk_order[bs.tb + rank] = i;
    }
  }
  __syncwarp();
  // slot assignment of straighten_formants @B35074 (sequential over the ranking)
  {
    double anchor = 0;
    int slot = 0;
    for (int r = 0; r < nr; r++) {
      const int i = p.trk_order[bs.tb + r];
      const double m = p.trk_mean[bs.tb + i];
      if (fabs(m - anchor) > 20) {
        anchor = m;
        slot++;
        if (slot >= 3) break;
      }
      if (lane == 0) p.trk_slot[bs.tb + i] = (signed char)slot;
    }
  }
  __syncwarp();

  const int si = st.n_segs;
  fa_segment seg;
  seg.start = start; seg.len = len; seg.stored = -1; seg.n_syllables = 0; seg.first_syllable = -1; seg.row_offset = -1;
  seg.ymax = st.y; seg.vmin = st.v; seg.cs_ratio = st.c_energy / st.s_energy;
  st.n_segs++;

  // rows: which points land on which frame row
  int* rc = p.row_count + bs.sb;
  int* ro = p.row_off + bs.sb;
  for (int r = lane; r < len; r += 32) rc[r] = 0;
  __syncwarp();
  bool thrown = false;
  for (int q = lane; q < st.n_pts; q += 32) {
    const int i = p.pt_track[bs.pb + q];
    if (p.trk_slot[bs.tb + i] >= 0) {
      const int fr = p.pt_frame[bs.pb + q];
      if (fr < 0 || fr >= len) thrown = true;  // r[d] is undefined -> TypeError -> .catch(L(-1))
      else atomicAdd(&rc[fr], 1);
    }
  }
  thrown = __any_sync(FULL, thrown);
  if (thrown) {
    if (lane == 0) p.segs[bs.sb + si] = seg;
    return -1;
  }
  __syncwarp();
  {
    int run = 0;
    for (int r0 = 0; r0 < len; r0 += 32) {
      const int r = r0 + lane;
      const int c = r < len ? rc[r] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
      }
      if (r < len) { ro[r] = run + incl - c; }
      run += __shfl_sync(FULL, incl, 31);
    }
  }
  __syncwarp();
  for (int r = lane; r < len; r += 32) rc[r] = 0;
  __syncwarp();
  for (int q = lane; q < st.n_pts; q += 32) {
    const int i = p.pt_track[bs.pb + q];
    if (p.trk_slot[bs.tb + i] >= 0) {
      const int fr = p.pt_frame[bs.pb + q];
      const int pos = ro[fr] + atomicAdd(&rc[fr], 1);
      p.row_list[bs.pb + pos] = q;
    }
  }
  __syncwarp();
  // apply: lane per row, points ordered by (track rank, point ordinal)
  float* Fout = p.formants + (size_t)(bs.row0 + st.n_rows) * 9;
  float* Eout = p.energy + (size_t)(bs.row0 + st.n_rows) * 3;
  for (int fr = lane; fr < len; fr += 32) {
    const int k = rc[fr], off = ro[fr];
    float f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0, f6 = 0, f7 = 0, f8 = 0, g0 = 0, g1 = 0, g2 = 0;
    long long last = -1;
    for (int it = 0; it < k; it++) {
      long long bestkey = 0x7fffffffffffffffll;
      int bq = -1;
      for (int z = 0; z < k; z++) {
        const int q = p.row_list[bs.pb + off + z];
        const int i = p.pt_track[bs.pb + q];
        const long long key = ((long long)p.trk_rank[bs.tb + i] << 32) | (unsigned)p.pt_ord[bs.pb + q];
        if (key > last && key < bestkey) { bestkey = key; bq = q; }
      }
      last = bestkey;
      const int i = p.pt_track[bs.pb + bq];
      int sl = p.trk_slot[bs.tb + i];
      const int bs_ = p.pt_binspan[bs.pb + bq];
      const int bin = bs_ & 0xffff, span = bs_ >> 16;
      const double E = p.pt_e[bs.pb + bq];
      const float cur = sl == 0 ? f0 : sl == 1 ? f3 : f6;
      if ((double)cur > vmin && (double)cur < (double)bin && sl < 2) sl++;
      const float fb = (float)bin, fe = (float)E, fs = (float)span;
      if (sl == 0) { f0 = fb; f1 = fe; f2 = fs; }
      else if (sl == 1) { f3 = fb; f4 = fe; f5 = fs; }
      else { f6 = fb; f7 = fe; f8 = fs; }
      g0 = (float)((double)g0 + (double)bin * E);
      g1 = (float)((double)g1 + E);
      g2 = (float)((double)g2 + (double)span * E);
    }
    float* fo = Fout + (size_t)fr * 9;
    fo[0] = f0; fo[1] = f1; fo[2] = f2; fo[3] = f3; fo[4] = f4; fo[5] = f5; fo[6] = f6; fo[7] = f7; fo[8] = f8;
    float* eo = Eout + (size_t)fr * 3;
    eo[0] = g0; eo[1] = g1; eo[2] = g2;
  }
  __syncwarp();
  seg.stored = st.n_stored;
  seg.row_offset = st.n_rows;
  seg.first_syllable = st.n_syls;
  // sep_syllables @B34757
  int nsyl = 0;
  if (p.level == 10 || p.level == 11 || p.level == 13) {
    int sstart = -1, quiet = 0, loud = 0;
    for (int e0 = 0; e0 < len; e0 += 32) {
      const int e = e0 + lane;
      const bool is_loud = e < len && (double)Eout[(size_t)e * 3 + 1] > vmin;
      const unsigned bits = __ballot_sync(FULL, is_loud);
      const int cnt = min(32, len - e0);
      for (int b = 0; b < cnt; b++) {
        const int ee = e0 + b;
        if ((bits >> b) & 1u) { quiet = 0; loud++; if (sstart < 0) sstart = ee; }
        else quiet++;
        if ((loud > 20 && quiet > 0) || (loud > 10 && quiet > 1) || (loud > 0 && quiet > 4) || (ee >= len - 1 && loud > 4)) {
          const int end = ee - quiet;
          if (end - sstart > 1) {
            if (lane == 0) {
              fa_syllable sy;
              sy.stored_seg = st.n_stored; sy.start = sstart; sy.len = end - sstart; sy.reserved = 0;
              p.syls[bs.sb + st.n_syls + nsyl] = sy;
            }
            nsyl++;
            sstart = -1;
            loud = 0;
          }
        }
      }
    }
  }
  seg.n_syllables = nsyl;
  if (lane == 0) p.segs[bs.sb + si] = seg;
  st.n_syls += nsyl;
  st.n_stored++;
  st.n_rows += len;
  return 1;
}

__global__ void __launch_bounds__(128) fa_segment_kernel(const FaSegmentParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpShared* sh = reinterpret_cast<WarpShared*>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int ui = blockIdx.x * (int)(blockDim.x >> 5) + wib;
  if (ui >= p.utt_count) return;
  const int u = p.utt_begin + ui;
  WarpShared& S = sh[wib];
  Bases bs;
  bs.row0 = p.frame_off[u];
  bs.F = (int)(p.frame_off[u + 1] - bs.row0);
  bs.tb = p.track_base[u];
  bs.tcap = (int)(p.track_base[u + 1] - bs.tb);
  bs.pb = bs.row0 * p.maxp;
  bs.sb = bs.row0 + u;
  const int B = p.B;

  ScanState st;
  st.current_frame = 0; st.no_fm_segs = 0; st.c_ci = 0; st.c_started = -1; st.w = 0; st.k = 0;
  st.y = p.y0; st.v = p.v0; st.x = p.y0; st.v0 = p.v0; st.T = 0; st.s_energy = 0; st.c_energy = 0;
  st.n_tr = 0; st.n_act = 0; st.n_pts = 0; st.n_segs = 0; st.n_stored = 0; st.n_rows = 0; st.n_syls = 0;
  st.overflow = 0;

  const int epl = (B + 31) >> 5;                     // frame words per lane (<= 8)
  const int cpl = (p.maxp + 31) >> 5;                // candidate words per lane (<= 5)
  uint32_t e_next[8], c_next[5];
  int nc_next = 0;
  double g_next = 0;
  auto prefetch = [&](int t) {
    const uint32_t* fr = p.frames + (size_t)(bs.row0 + t) * B;
    const uint32_t* cd = p.cand + (size_t)(bs.row0 + t) * p.maxp;
#pragma unroll
    for (int i = 0; i < 8; i++) e_next[i] = (i < epl && lane + 32 * i < B) ? __ldg(fr + lane + 32 * i) : 0u;
#pragma unroll
    for (int i = 0; i < 5; i++) c_next[i] = (i < cpl && lane + 32 * i < p.maxp) ? __ldg(cd + lane + 32 * i) : 0u;
    nc_next = __ldg(p.ncand + bs.row0 + t);
    g_next = __ldg(p.gsum + bs.row0 + t);
  };
  if (bs.F > 0) prefetch(0);

  for (int t = 0; t < bs.F && !st.overflow; t++) {
    // ---- spectrum_push @B30392 ----
    st.current_frame++;
    uint32_t cand[5];
#pragma unroll
    for (int i = 0; i < 8; i++) if (i < epl && lane + 32 * i < B) S.e[lane + 32 * i] = e_next[i];
#pragma unroll
    for (int i = 0; i < 5; i++) cand[i] = c_next[i];
    const int nc = min(nc_next, p.maxp);
    if (nc_next > p.maxp) st.overflow = 1;
    const double g = g_next;
    __syncwarp();
    if (t + 1 < bs.F) prefetch(t + 1);

    // ---- D() @B25717: filter the candidates of K2 by the gate v (value at frame start) ----
    const double v = st.v;
    const int t_stale = st.c_ci;
    int n = 0, pbin = 0;
    unsigned long long dsum = 0;
    double h = 2 * v;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      if (i * 32 < nc) {
        const int c = i * 32 + lane;
        const bool valid = c < nc;
        const uint32_t pkd = cand[i];
        const int pk = (pkd >> 16) & 0xff;
        const uint32_t amp = valid ? S.e[pk] : 0u;
        const bool acc = valid && (double)amp > v;
        const unsigned m = __ballot_sync(FULL, acc);
        if (acc) {
          const int pos = n + __popc(m & ((1u << lane) - 1));
          if (pos < PCAP) { S.plo[pos] = pkd & 0xff; S.phi[pos] = (pkd >> 8) & 0xff; S.ppk[pos] = pk; }
        }
        n += __popc(m);
        dsum += warp_sum_u64(acc ? (unsigned long long)amp : 0ull);
        // h / p: first strictly greater wins; the last-bin peak never updates them
        const bool hp = acc && !((pkd >> 24) & 1u);
        uint32_t mx = hp ? amp : 0u;
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
        const unsigned who = __ballot_sync(FULL, hp && amp == mx);
        if (who && (double)mx > h) {
          h = (double)mx;
          pbin = __shfl_sync(FULL, pk, __ffs(who) - 1);
        }
      }
    }
    if (n > PCAP) { st.overflow = 1; break; }
    __syncwarp();
    const double d = (double)dsum;

    int fin = -2;
    if (st.c_started < 0) {
      const double ratio = d > h ? h * (double)(n - 1) / (d - h) : 0;
      if (n > 0 && pbin > 7 && pbin < p.max_voiced_bin && n > 4 && ratio > 4) { seg_reset(st, 0); st.c_started = 0; }
      else st.no_fm_segs++;
    }
    if (st.c_started >= 0) {
      if (n == 0 || pbin < 7 || pbin >= p.max_voiced_bin || (n > 3 && d / (g - d) < 0.1)) {
        st.no_fm_segs++;
        if (st.c_started < 2) st.c_started--;
        else if ((double)st.no_fm_segs >= p.seg_breaker) fin = finalize_segment(p, S, st, bs, st.c_ci + 1, lane);
        else if (p.auto_gate) noise_gate(st, h);
      } else {
        if (p.auto_gate) noise_gate(st, h);
        accumulate_fm(p, S, st, bs, n, t_stale, g, st.v, lane);
        if (st.c_started < 2) st.c_started++; else st.no_fm_segs = 0;
      }
    }
    st.c_ci++;
    if (fin != -2) seg_reset(st, -1);  // the promise's micro-task runs before the next frame
    __syncwarp();
  }
  // segment_truncate @B30800
  if (!st.overflow) {
    finalize_segment(p, S, st, bs, st.c_ci, lane);
    seg_reset(st, 1);
  }
  if (lane == 0) {
    p.n_segs[u] = st.n_segs;
    p.n_stored[u] = st.n_stored;
    p.n_rows[u] = st.n_rows;
    p.n_syls[u] = st.n_syls;
    p.overflow[u] = st.overflow;
  }
}

}  // namespace

cudaError_t fa_launch_segment(const FaSegmentParams& p, cudaStream_t s, int* launches) {
  if (p.utt_count <= 0) return cudaSuccess;
  int kw = kWarps;
  if (const char* ev = getenv("FA_K3_WARPS")) { const int v = atoi(ev); if (v >= 1 && v <= 4) kw = v; }  // tuning knob
  const int grid = (p.utt_count + kw - 1) / kw;
  const int bytes = (int)sizeof(WarpShared) * kw;
  cudaError_t e = cudaFuncSetAttribute(fa_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  fa_segment_kernel<<<grid, kw * 32, bytes, s>>>(p);
  if (launches) (*launches)++;
  return cudaGetLastError();
}
