// fa_synth.cpp -- synthetic "glottal pulse through formant resonators" speech (host code).
//
// The workload generator of SURVEY.md section 8(d) / BASELINE.json ("synthetic glottal-pulse-through-
// formant-resonator speech"): a Rosenberg glottal pulse train (open quotient 0.6, differentiated) with
// F0 ~ U(90,250) Hz, +-3 % vibrato at 5 Hz and 1 % jitter, through a cascade of four two-pole
// resonators (F1 ~ U(300,900), F2 ~ U(900,2300), F3 ~ U(2300,3200), F4 = 3500 Hz; bandwidths
// 60/90/120/150 Hz) re-drawn every syllable; raised-cosine syllable bursts of U(120,350) ms separated
// by U(30,80) ms gaps, phrases of 3-8 syllables separated by U(250,600) ms silences; white noise at
// -50 dBFS; peak-normalised to 0.3.  Deterministic in (seed, utt_index); no reference counterpart
// (the reference ships one demo WAV and no generator).
// Built as its own small host library (libfa_synth.so): workload generation is not part of the product library, so a process
// that only times a CPU baseline never maps libfa_b200.so.
#include <cmath>
#include <cstdint>
#include <vector>

#include "fa_synth.h"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  double uni(double a, double b) { return a + (b - a) * uni(); }
};

struct Resonator {
  double a1 = 0, a2 = 0, b0 = 1, y1 = 0, y2 = 0;
  void set(double f, double bw, double sr) {
    const double r = std::exp(-M_PI * bw / sr), th = 2.0 * M_PI * f / sr;
    a1 = 2.0 * r * std::cos(th);
    a2 = -r * r;
    b0 = 1.0 - a1 - a2;
  }
  double step(double x) {
    const double y = b0 * x + a1 * y1 + a2 * y2;
    y2 = y1;
    y1 = y;
    return y;
  }
};

double rosenberg(double ph) {
  const double tp = 0.4, tn = 0.2;
  if (ph < tp) return 0.5 * (1.0 - std::cos(M_PI * ph / tp));
  if (ph < tp + tn) return std::cos(M_PI * (ph - tp) / (2.0 * tn));
  return 0.0;
}

}  // namespace

extern "C" int fa_synth_speech(float* dst, size_t n, int sr, uint64_t seed, uint64_t utt) {
  if (!dst || sr < 4000) return -1;
  Rng rng(0x5EEDull ^ seed ^ (utt * 0xD1B54A32D192ED03ull));
  std::vector<double> x(n, 0.0);
  const double nyq = 0.5 * sr;
  Resonator res[4];
  size_t pos = (size_t)(rng.uni(0.02, 0.15) * sr);
  double phase = 0.0, prev_g = 0.0;
  while (pos < n) {
    const int n_syl = 3 + (int)(rng.uni() * 6.0);  // 3..8
    const double f0_base = rng.uni(90.0, 250.0);
    for (int sidx = 0; sidx < n_syl && pos < n; sidx++) {
      const size_t len = (size_t)(rng.uni(0.120, 0.350) * sr);
      double f[4] = {rng.uni(300, 900), rng.uni(900, 2300), rng.uni(2300, 3200), 3500.0};
      const double bw[4] = {60, 90, 120, 150};
      for (int k = 0; k < 4; k++) {
        if (f[k] > 0.9 * nyq) f[k] = 0.9 * nyq;
        res[k].set(f[k], bw[k], sr);
      }
      const double f0_syl = f0_base * rng.uni(0.9, 1.1);
      const double amp = rng.uni(0.5, 1.0);
      double jitter = 1.0;
      for (size_t i = 0; i < len && pos + i < n; i++) {
        const double t = (double)(pos + i) / sr;
        const double f0 = f0_syl * (1.0 + 0.03 * std::sin(2.0 * M_PI * 5.0 * t)) * jitter;
        phase += f0 / sr;
        if (phase >= 1.0) { phase -= 1.0; jitter = 1.0 + 0.01 * (2.0 * rng.uni() - 1.0); }
        const double g = rosenberg(phase);
        double v = g - prev_g;  // lip radiation
        prev_g = g;
        for (int k = 0; k < 4; k++) v = res[k].step(v);
        const double env = 0.5 * (1.0 - std::cos(2.0 * M_PI * (double)i / (double)len));
        x[pos + i] = amp * env * v;
      }
      pos += len + (size_t)(rng.uni(0.030, 0.080) * sr);
    }
    pos += (size_t)(rng.uni(0.250, 0.600) * sr);
  }
  double peak = 1e-12;
  for (size_t i = 0; i < n; i++) peak = std::fmax(peak, std::fabs(x[i]));
  const double scale = 0.3 / peak, noise = std::pow(10.0, -50.0 / 20.0) * std::sqrt(3.0);  // uniform noise, rms -50 dBFS
  for (size_t i = 0; i < n; i++) dst[i] = (float)(x[i] * scale + noise * (2.0 * rng.uni() - 1.0));
  return 0;
}

// A whole batch as 16-bit PCM (what a WAV file holds): utterance i = index first_index + i * index_stride, n samples each,
// written back to back; rint(x * 32768) clipped.  OpenMP over utterances.
extern "C" int fa_synth_speech_i16_batch(int16_t* dst, int n_utt, size_t n, int sr, uint64_t seed, uint64_t first_index,
                                         uint64_t index_stride, int threads) {
  if (!dst || n_utt < 0 || sr < 4000) return -1;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1) reduction(+ : bad)
  for (int i = 0; i < n_utt; i++) {
    std::vector<float> x(n);
    if (fa_synth_speech(x.data(), n, sr, seed, first_index + (uint64_t)i * index_stride) != 0) { bad++; continue; }
    int16_t* o = dst + (size_t)i * n;
    for (size_t k = 0; k < n; k++) {
      double v = std::nearbyint((double)x[k] * 32768.0);
      v = v < -32768.0 ? -32768.0 : v > 32767.0 ? 32767.0 : v;
      o[k] = (int16_t)v;
    }
  }
  return bad ? -1 : 0;
}
