'use strict';
/*
 * index.js -- drop-in for `require('formantanalyzer')` on Node.js, backed by the B200 addon (fa_b200.node ->
 * libfa_b200.so).  Same four exports, argument order, callback shapes and string rejections as the reference's
 * inner module 1 (/root/reference/dist/main.js line 2: configure @B3292, LaunchAudioNodes @B4469, StopAudioNodes
 * @B5699, set_predicted_label_for_segment).  Python twin: webspeechanalyzer_b200/api.py (tested on the GPU box).
 *
 * Node has no AudioContext: context_source 1 takes a WAV ArrayBuffer/Buffer (RIFF PCM 8/16/24/32 or float32/64),
 * the extension context_source 4 takes {pcm: Float32Array, sampleRate}; 2 (<audio>) and 3 (mic) reject with
 * "Invalid audio source".  Levels 1-2 (canvas only in the reference) call back once with the dB spectrum.
 *
 * Status: written against the C-ABI and compile-checked; no Node in the build image, so the file as a whole has not run --
 * but its two pure functions have: configure() and segmentCallbacks() are executed by the test suite's JavaScript interpreter in
 * tests/test_host_api.py (against the reference's own configure() and against the Python twin at every output level).
 */
const native = require('./fa_b200.node');

const DEFAULTS = Object.freeze({ // @B2972
  plot_enable: false, spec_type: 1, output_level: 4, plot_len: 200, f_min: 50, f_max: 4000, N_fft_bins: 256,
  N_mel_bins: 128, window_width: 25, window_step: 25, pause_length: 200, min_seg_length: 50, auto_noise_gate: true,
  voiced_max_dB: 100, voiced_min_dB: 10, plot_lag: 1, pre_norm_gain: 1000, high_f_emph: 0, plot_canvas: null,
  canvas_width: 200, canvas_height: 100,
  fftSize: 2048, smoothingTimeConstant: 0.8, minDecibels: -100, maxDecibels: -30, mag_scale: 0, clamp_dB: true, device: 0,
});
let a = Object.assign({}, DEFAULTS);
const state = { playing: false, stop: null, engine: null, engineKey: '', labels: [] };

function configure(e) { // truthiness rules of @B3292
  if (e.spec_type !== null && e.spec_type !== undefined) a.spec_type = e.spec_type;
  if (e.output_level) a.output_level = e.output_level;
  if (e.f_min !== null && e.f_min !== undefined) a.f_min = e.f_min;
  if (e.f_max) a.f_max = e.f_max;
  if (e.N_fft_bins) a.N_fft_bins = e.N_fft_bins;
  if (e.N_mel_bins) a.N_mel_bins = e.N_mel_bins;
  if (e.window_width) a.window_width = e.window_width;
  if (e.window_step) a.window_step = e.window_step;
  if (e.pre_norm_gain) a.pre_norm_gain = e.pre_norm_gain;
  if (e.high_f_emph !== null && e.high_f_emph !== undefined) a.high_f_emph = e.high_f_emph;
  if (e.pause_length) a.pause_length = e.pause_length;
  if (e.min_seg_length) a.min_seg_length = e.min_seg_length;
  if (e.auto_noise_gate !== null && e.auto_noise_gate !== undefined) a.auto_noise_gate = e.auto_noise_gate;
  if (e.voiced_max_dB) a.voiced_max_dB = e.voiced_max_dB;
  if (e.voiced_min_dB !== null && e.voiced_min_dB !== undefined) a.voiced_min_dB = e.voiced_min_dB;
  if (e.plot_enable && e.plot_canvas) { a.plot_enable = true; if (e.plot_len) a.plot_len = e.plot_len; } else a.plot_enable = false;
  for (const k of ['fftSize', 'smoothingTimeConstant', 'minDecibels', 'maxDecibels', 'mag_scale', 'clamp_dB', 'device'])
    if (e[k] !== null && e[k] !== undefined) a[k] = e[k];
}

function decodeWav(buf) { // stand-in for decodeAudioData (@B18693); no resampling (SURVEY.md 8, sample-rate caveat)
  const u8 = buf instanceof ArrayBuffer ? new Uint8Array(buf) : new Uint8Array(buf.buffer, buf.byteOffset, buf.byteLength);
  const dv = new DataView(u8.buffer, u8.byteOffset, u8.byteLength);
  const tag4 = (o) => String.fromCharCode(u8[o], u8[o + 1], u8[o + 2], u8[o + 3]);
  if (u8.length < 12 || tag4(0) !== 'RIFF' || tag4(8) !== 'WAVE') throw 'Unable to decode audio data';
  let pos = 12, fmt = null, data = null;
  while (pos + 8 <= u8.length) {
    const id = tag4(pos), size = dv.getUint32(pos + 4, true);
    if (id === 'fmt ') {
      let tag = dv.getUint16(pos + 8, true);
      if (tag === 0xFFFE && size >= 26) tag = dv.getUint16(pos + 8 + 24, true);
      fmt = { tag, ch: dv.getUint16(pos + 10, true), sr: dv.getUint32(pos + 12, true), bits: dv.getUint16(pos + 22, true) };
    } else if (id === 'data') data = { off: pos + 8, size: Math.min(size, u8.length - pos - 8) };
    pos += 8 + size + (size & 1);
  }
  if (!fmt || !data || fmt.ch < 1 || fmt.sr < 1) throw 'Unable to decode audio data';
  // linear PCM (1) and IEEE float (3) only: A-law / mu-law / ADPCM data must not be read as linear samples (wav.py: same rule)
  if (fmt.tag !== 1 && fmt.tag !== 3) throw 'Unable to decode audio data';
  if (fmt.tag === 3 && fmt.bits !== 32 && fmt.bits !== 64) throw 'Unable to decode audio data';
  const bps = fmt.bits / 8, n = Math.floor(data.size / (bps * fmt.ch)), out = new Float32Array(n);
  for (let i = 0; i < n; i++) {
    let acc = 0;
    for (let c = 0; c < fmt.ch; c++) {
      const o = data.off + (i * fmt.ch + c) * bps;
      let v;
      if (fmt.tag === 3) v = fmt.bits === 32 ? dv.getFloat32(o, true) : dv.getFloat64(o, true);
      else if (fmt.bits === 8) v = (u8[o] - 128) / 128;
      else if (fmt.bits === 16) v = dv.getInt16(o, true) / 32768;
      else if (fmt.bits === 24) { let x = u8[o] | (u8[o + 1] << 8) | (u8[o + 2] << 16); if (x & 0x800000) x -= 0x1000000; v = x / 8388608; }
      else if (fmt.bits === 32) v = dv.getInt32(o, true) / 2147483648;
      else throw 'Unable to decode audio data';
      acc += v;
    }
    out[i] = acc / fmt.ch;
  }
  return { pcm: out, sampleRate: fmt.sr };
}

function engine() {
  const key = JSON.stringify(a);
  if (!state.engine || state.engineKey !== key) {
    if (state.engine) native.destroyEngine(state.engine);   // frees the old handle's GPU memory now, not at the next GC
    state.engine = native.createEngine(Object.assign({}, a, { want_spectrum: a.output_level <= 2 ? 1 : 0 }), a.device | 0);
    state.engineKey = key;
  }
  return state.engine;
}

// P() @B28869: one callback per stored segment, time stamps from seg_ci[e] (get_seg_timestamps @B31504,
// get_syls_timestamps @B31114 -> toFixed(3) strings).  Store e is stamped with seg_ci[e] even after a dropped
// segment, like the reference (DESIGN.md quirk 15).
function segmentCallbacks(level, res, labels) {
  const step = a.window_step / 1e3, calls = [];
  const stored = res.segments.filter((s) => s.stored >= 0);
  if (level === 11) {
    // b(0, label, get_clip_timestamps() @B31365, get_utterance_features(u, h) @B107983) after every stored segment
    let k = 0;
    res.segments.forEach((s, si) => {
      if (s.stored < 0) return;
      let t = 0;
      for (let n = 0; n <= si; n++) t += res.segments[n].len;
      calls.push([0, labels, [res.segments[0].start * step, (t + 1) * step], Array.from(res.features.subarray(264 * k, 264 * (k + 1)))]);
      k++;
    });
    return calls;
  }
  if (level === 3) {
    // b(e, label, s[e]) when s[e].length > 0 (P() @B28869, level-3 branch): the ranked 18-field track arrays of
    // accumulate_fm @B35952, rebuilt from the fa_track headers (in `syllables`) and the six point columns
    const P = res.trackPoints, np = P.length / 6;
    stored.forEach((s, e) => {
      const tracks = [];
      for (let k = 0; k < s.nSyllables; k++) {
        const t = res.syllables[s.firstSyllable + k];
        const col = (c) => Array.from(P.subarray(c * np + t.start, c * np + t.start + t.len));
        const fr = col(0), lo = col(1), hi = col(2), bins = col(3), amp = col(4), en = col(5);
        const h = bins.length - 1, o = bins[h];
        let vel = 0;
        if (h >= 3) vel = (o - bins[h - 1] + (bins[h - 2] - bins[h - 1]) + (bins[h - 3] - bins[h - 2])) / 3;
        else if (h === 2) vel = (o - bins[h - 1] + (bins[h - 2] - bins[h - 1])) / 2;
        else if (h === 1) vel = o - bins[h - 1];
        let se = 0, seb = 0, span = 0;
        for (let i = 0; i <= h; i++) { se += en[i]; seb += en[i] * bins[i]; span += hi[i] - lo[i] + 1; }
        tracks.push([lo[h], hi[h], fr[h], fr[h], vel, o, amp[h], fr, lo, hi, bins, amp, en, se, h + 1, seb, 0, span]);
      }
      if (tracks.length) calls.push([e, labels, tracks]);
    });
    return calls;
  }
  stored.forEach((s, e) => {
    const ci = res.segments[e];
    const rows = [];
    for (let r = 0; r < s.len; r++) rows.push(res.formants.subarray(9 * (s.rowOffset + r), 9 * (s.rowOffset + r + 1)));
    if (level === 13 || level === 12 || level === 10) {
      const syl = res.syllables.slice(s.firstSyllable, s.firstSyllable + s.nSyllables);
      if (!syl.length) return;
      if (level === 12) {
        // make_coeffs @B34527 returns the rows made before numeric threw; P() fires when there is at least one, with the time
        // stamps of ALL the segment's syllables
        let ok = syl.findIndex((y) => y.flag);
        if (ok < 0) ok = syl.length;
        if (!ok) return;
        const t12 = syl.map((y) => [((ci.start + y.start) * step).toFixed(3), ((y.len + 1) * step).toFixed(3)]);
        const rows12 = [];
        for (let k = 0; k < ok; k++) rows12.push(Array.from(res.features.subarray(23 * (s.firstSyllable + k), 23 * (s.firstSyllable + k + 1))));
        calls.push([e, labels, t12, rows12]);
        return;
      }
      const times = syl.map((y) => [((ci.start + y.start) * step).toFixed(3), ((y.len + 1) * step).toFixed(3)]);
      const payload = level === 13
        ? syl.map((_, k) => Array.from(res.features.subarray(53 * (s.firstSyllable + k), 53 * (s.firstSyllable + k + 1))))
        : syl.map((y) => rows.slice(y.start, y.start + y.len));
      calls.push([e, labels, times, payload]);
    } else if (level === 5) {
      calls.push([e, labels, [ci.start * step, (ci.len + 1) * step], Array.from(res.features.subarray(53 * e, 53 * (e + 1)))]);
    } else if (level === 4 && rows.length) {
      calls.push([e, labels, [ci.start * step, (ci.len + 1) * step], rows]);
    }
  });
  return calls;
}

function LaunchAudioNodes(context_source, source_obj = null, callback = null, file_labels = [], offline = false,
  test_play = true, play_offset = null, play_duration = null) {
  return new Promise((resolve, reject) => {
    if (state.playing) return reject('Error: Already playing');
    if (!(a.N_mel_bins | a.N_fft_bins)) return reject('Invalid reset_nodes config');
    if (!(a.spec_type === 1 ? a.N_mel_bins : a.N_fft_bins)) return reject('reset_segmentor failed: Invalid spec_bands');
    if (!a.output_level) return reject('Invalid reset_plot config');
    let src;
    try {
      if (context_source === 1 && source_obj) src = decodeWav(source_obj);
      else if (context_source === 4 && source_obj) src = { pcm: source_obj.pcm, sampleRate: source_obj.sampleRate | 0 };
      else return reject('Invalid audio source');
    } catch (e) { return reject(e); }
    let pcm = src.pcm;
    if (play_offset || play_duration) {
      const o = Math.round((play_offset || 0) * src.sampleRate);
      const n = play_duration ? Math.round(play_duration * src.sampleRate) : pcm.length - o;
      pcm = pcm.subarray(o, o + Math.max(n, 0));
    }
    state.playing = true;
    state.labels = [];
    const level = a.output_level;
    let eng;
    try { eng = engine(); } catch (e) { state.playing = false; return reject(String(e.message || e)); }
    native.analyze(eng, pcm, src.sampleRate, level <= 2, a.fftSize, level).then((res) => {
      if (level <= 2) {
        if (!test_play && callback) callback(0, file_labels, [0, res.counts.frames * a.window_step / 1e3], res.spectrum);
      } else {
        const calls = segmentCallbacks(level, res, file_labels);
        state.labels = calls.map(() => file_labels.slice());
        if (!test_play && callback) for (const c of calls) { if (state.stop !== null) break; callback(...c); }
      }
      state.playing = false; state.stop = null;
      resolve(true);
    }).catch((e) => { state.playing = false; state.stop = null; reject(e); });
  });
}

function StopAudioNodes(reason = 'no reason') { if (state.playing) state.stop = reason; } // disconnect_nodes @B21559

function set_predicted_label_for_segment(seg_index, label_index, predicted_label) { // set_segments_label (module 3 `G`)
  const f = state.labels[seg_index];
  while (f.length < label_index) f[f.length] = -1;
  f[label_index] = predicted_label;
}

// Incremental callbacks for audio that is still arriving (the reference's microphone / <audio> sources): the segmentor is
// causal, so what a PREFIX of the stream finalises without segment_truncate @B30800 is what the reference has called back by
// then.  push() appends a chunk, re-analyses the stream so far with truncation off (> 300 000 x real time) and fires the new
// callbacks; stop() runs once more with truncation on.  Python twin (tested on the GPU): api.LiveSession.
function createLiveSession(sampleRate, callback, file_labels = []) {
  if (a.output_level <= 2) throw new Error('LiveSession delivers segment callbacks: output_level must be >= 3');
  const eng = engine(), level = a.output_level;
  let pcm = new Float32Array(0), fired = 0, chain = Promise.resolve(0);
  const run = (truncate) => {
    native.setTruncate(eng, truncate);
    return native.analyze(eng, pcm, sampleRate, false, a.fftSize, level).then((res) => {
      native.setTruncate(eng, true);
      const calls = segmentCallbacks(level, res, file_labels), fresh = calls.slice(fired);
      fired = calls.length;
      if (callback) for (const c of fresh) callback(...c);
      return fresh.length;
    });
  };
  return {
    push(chunk) {
      chain = chain.then(() => {
        const grown = new Float32Array(pcm.length + chunk.length);
        grown.set(pcm); grown.set(chunk, pcm.length);
        pcm = grown;
        return run(false);
      });
      return chain;
    },
    stop() { chain = chain.then(() => run(true)); return chain; },
  };
}

module.exports = { configure, LaunchAudioNodes, StopAudioNodes, set_predicted_label_for_segment, createLiveSession };
