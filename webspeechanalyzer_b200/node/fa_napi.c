/*
 * fa_napi.c -- Node.js addon (N-API) over the C-ABI of libfa_b200.so (include/fa_b200.h).
 *
 * The reference's host language is JavaScript: formantanalyzer is `require`d by the app
 * (/root/reference/src/index.js:9) and exposes configure / LaunchAudioNodes / StopAudioNodes /
 * set_predicted_label_for_segment (/root/reference/dist/main.js:2@B2750).  index.js in this directory keeps
 * that surface; this file is the thin native part it calls:
 *
 *   createEngine(cfg, device) -> external handle          fa_create
 *   destroyEngine(handle)                                  fa_destroy
 *   analyze(handle, Float32Array pcm, sampleRate) -> Promise<{counts, segments, formants, energy, syllables,
 *            features, spectrum?}>                         fa_reset + fa_submit_pcm + fa_run (+ fa_sync, fa_copy_*)
 *
 * The JS thread never blocks: the CUDA work runs in napi_async_work's execute callback (libuv pool thread);
 * results are copied into JS-owned ArrayBuffers in the complete callback, back on the JS thread.
 * Status: compile-checked in this image (no Node.js here); loaded and tested only where `node` exists.
 */
#ifdef FA_USE_SYSTEM_NODE_API
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fa_b200.h"

#define CHECK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, "N-API call failed: " #call); return NULL; } } while (0)

static double get_num(napi_env env, napi_value obj, const char* key, double dflt) {
  bool has = false;
  napi_value v;
  napi_valuetype t;
  double d;
  bool b;
  if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return dflt;
  if (napi_get_named_property(env, obj, key, &v) != napi_ok || napi_typeof(env, v, &t) != napi_ok) return dflt;
  if (t == napi_number && napi_get_value_double(env, v, &d) == napi_ok) return d;
  if (t == napi_boolean && napi_get_value_bool(env, v, &b) == napi_ok) return b ? 1.0 : 0.0;
  return dflt;
}

/* the external wraps a box, so that destroyEngine can free the GPU memory at once and leave the finalizer nothing to do */
typedef struct { fa_handle* h; } engine_box;

static void finalize_engine(napi_env env, void* data, void* hint) {
  (void)env; (void)hint;
  engine_box* b = (engine_box*)data;
  if (!b) return;
  if (b->h) fa_destroy(b->h);
  free(b);
}

/* createEngine(cfg, device): cfg uses formantanalyzer's field names (defaults @B2972) + the AnalyserNode extension */
static napi_value CreateEngine(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2], out;
  CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  fa_config c;
  fa_config_default(&c);
  int32_t device = 0;
  if (argc >= 1) {
    napi_value o = argv[0];
    c.spec_type = (int32_t)get_num(env, o, "spec_type", c.spec_type);
    c.output_level = (int32_t)get_num(env, o, "output_level", c.output_level);
    c.plot_len = (int32_t)get_num(env, o, "plot_len", c.plot_len);
    c.n_fft_bins = (int32_t)get_num(env, o, "N_fft_bins", c.n_fft_bins);
    c.n_mel_bins = (int32_t)get_num(env, o, "N_mel_bins", c.n_mel_bins);
    c.auto_noise_gate = (int32_t)get_num(env, o, "auto_noise_gate", c.auto_noise_gate);
    c.f_min = get_num(env, o, "f_min", c.f_min);
    c.f_max = get_num(env, o, "f_max", c.f_max);
    c.window_width_ms = get_num(env, o, "window_width", c.window_width_ms);
    c.window_step_ms = get_num(env, o, "window_step", c.window_step_ms);
    c.pause_length_ms = get_num(env, o, "pause_length", c.pause_length_ms);
    c.min_seg_length_ms = get_num(env, o, "min_seg_length", c.min_seg_length_ms);
    c.voiced_max_db = get_num(env, o, "voiced_max_dB", c.voiced_max_db);
    c.voiced_min_db = get_num(env, o, "voiced_min_dB", c.voiced_min_db);
    c.pre_norm_gain = get_num(env, o, "pre_norm_gain", c.pre_norm_gain);
    c.high_f_emph = get_num(env, o, "high_f_emph", c.high_f_emph);
    c.fft_size = (int32_t)get_num(env, o, "fftSize", c.fft_size);
    c.smoothing = get_num(env, o, "smoothingTimeConstant", c.smoothing);
    c.min_db = get_num(env, o, "minDecibels", c.min_db);
    c.max_db = get_num(env, o, "maxDecibels", c.max_db);
    c.mag_scale = get_num(env, o, "mag_scale", c.mag_scale);
    c.clamp_db = (int32_t)get_num(env, o, "clamp_dB", c.clamp_db);
    c.want_spectrum = (int32_t)get_num(env, o, "want_spectrum", c.want_spectrum);
  }
  if (argc >= 2) napi_get_value_int32(env, argv[1], &device);
  fa_handle* h = NULL;
  const int rc = fa_create(&c, device, &h);
  if (rc != FA_OK) {
    /* the strings the reference rejects with (@B4469 chain) where they apply */
    napi_throw_error(env, NULL, rc == FA_ERR_INVALID_ARG ? "Invalid reset_nodes config" : fa_status_string(rc));
    return NULL;
  }
  engine_box* box = (engine_box*)malloc(sizeof(engine_box));
  if (!box) { fa_destroy(h); napi_throw_error(env, NULL, "out of memory"); return NULL; }
  box->h = h;
  if (napi_create_external(env, box, finalize_engine, NULL, &out) != napi_ok) {
    fa_destroy(h); free(box);
    napi_throw_error(env, NULL, "N-API call failed: napi_create_external");
    return NULL;
  }
  return out;
}

/* destroyEngine(handle): release the handle's device and pinned memory now (StopAudioNodes / a new configuration) */
static napi_value DestroyEngine(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  void* bp = NULL;
  if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < 1) return NULL;
  if (napi_get_value_external(env, argv[0], &bp) != napi_ok || !bp) return NULL;
  engine_box* b = (engine_box*)bp;
  if (b->h) { fa_destroy(b->h); b->h = NULL; }   /* the finalizer only frees the box afterwards */
  return NULL;
}

/* setTruncate(handle, on): fa_set_truncate -- off while the analysed audio is the prefix of a stream that is still running */
static napi_value SetTruncate(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  void* bp = NULL;
  bool on = true;
  if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < 2) return NULL;
  if (napi_get_value_external(env, argv[0], &bp) != napi_ok || !bp || !((engine_box*)bp)->h) return NULL;
  if (napi_get_value_bool(env, argv[1], &on) != napi_ok) return NULL;
  fa_set_truncate(((engine_box*)bp)->h, on ? 1 : 0);
  return NULL;
}

typedef struct {
  napi_async_work work;
  napi_deferred deferred;
  napi_ref pcm_ref;
  fa_handle* h;
  const float* pcm;
  size_t n;
  int sr, rc, want_spec, fft_half, feat_width, is_frames;
  char err[256];
  fa_counts counts;
  fa_segment* segs;
  fa_syllable* syls;
  float *formants, *energy, *spectrum;
  double* features;
  fa_track_point* points;   /* level 3: the points of the ranked tracks (their headers come in `syls`) */
  int level;
} job_t;

static void job_execute(napi_env env, void* data) {
  (void)env;
  job_t* j = (job_t*)data;
  fa_handle* h = j->h;
  int rc = fa_reset(h);
  if (rc == FA_OK)
    rc = j->is_frames ? (j->sr > 0 ? fa_submit_frames(h, 0, (const uint32_t*)j->pcm, j->n / (size_t)j->sr, j->sr) : FA_ERR_INVALID_ARG)
                      : fa_submit_pcm(h, 0, j->pcm, j->n, j->sr);
  if (rc >= 0) rc = fa_run(h);
  if (rc == FA_OK) rc = fa_sync(h);
  if (rc == FA_OK) rc = fa_result_counts(h, 0, &j->counts);
  if (rc == FA_OK) {
    const fa_counts* c = &j->counts;
    /* row widths come from the handle's own configuration, never from what the caller said */
    fa_config cfg;
    fa_get_config(h, &cfg);
    j->fft_half = cfg.fft_size / 2;
    j->level = cfg.output_level;
    j->feat_width = cfg.output_level == FA_LEVEL_UTTERANCE ? FA_N_UTT_FEATURES
                    : cfg.output_level == FA_LEVEL_SYL_CURVES ? FA_N_CURVE_FEATURES : FA_N_FEATURES;
    j->want_spec = j->want_spec && !j->is_frames && (cfg.want_spectrum || cfg.output_level <= 2);
    j->segs = (fa_segment*)malloc(sizeof(fa_segment) * (size_t)(c->segments + 1));
    j->syls = (fa_syllable*)malloc(sizeof(fa_syllable) * (size_t)(c->syllables + 1));
    j->formants = (float*)malloc(sizeof(float) * 9 * (size_t)(c->formant_rows + 1));
    j->energy = (float*)malloc(sizeof(float) * 3 * (size_t)(c->formant_rows + 1));
    /* level 11 rows are the 264-dim utterance distributions (cumulative, one per stored segment) */
    j->features = (double*)malloc(sizeof(double) * (size_t)j->feat_width * (size_t)(c->feature_rows + 1));
    if (j->want_spec) j->spectrum = (float*)malloc(sizeof(float) * (size_t)j->fft_half * (size_t)(c->frames + 1));
    if (j->level == FA_LEVEL_SEGMENTS) j->points = (fa_track_point*)malloc(sizeof(fa_track_point) * (size_t)(c->formant_rows + 1));
    if (!j->segs || !j->syls || !j->formants || !j->energy || !j->features || (j->want_spec && !j->spectrum) ||
        (j->level == FA_LEVEL_SEGMENTS && !j->points)) {
      j->rc = FA_ERR_OUT_OF_MEMORY;
      snprintf(j->err, sizeof(j->err), "out of memory");
      return;
    }
    rc = fa_copy_segments(h, 0, j->segs, (size_t)c->segments);
    if (rc >= 0) rc = fa_copy_syllables(h, 0, j->syls, (size_t)c->syllables);
    if (j->level == FA_LEVEL_SEGMENTS) {   /* raw ranked tracks: fa_track headers in `syls`, their points here; no formant rows */
      if (rc >= 0) rc = fa_copy_track_points(h, 0, j->points, (size_t)c->formant_rows);
    } else {
      if (rc >= 0) rc = fa_copy_formants(h, 0, j->formants, (size_t)c->formant_rows);
      if (rc >= 0) rc = fa_copy_energy(h, 0, j->energy, (size_t)c->formant_rows);
    }
    if (rc >= 0) rc = j->feat_width == FA_N_UTT_FEATURES ? fa_copy_utterance_features(h, 0, j->features, (size_t)c->feature_rows)
                  : j->feat_width == FA_N_CURVE_FEATURES ? fa_copy_curve_features(h, 0, j->features, (size_t)c->feature_rows)
                                                         : fa_copy_features(h, 0, j->features, (size_t)c->feature_rows);
    if (rc >= 0 && j->want_spec) rc = fa_copy_spectrum(h, 0, j->spectrum, (size_t)c->frames);
    if (rc >= 0) rc = FA_OK;
  }
  j->rc = rc;
  if (rc != FA_OK) snprintf(j->err, sizeof(j->err), "%s", fa_last_error(h));
}

static napi_value make_f32(napi_env env, const float* src, size_t n) {
  void* data = NULL;
  napi_value ab, ta;
  if (napi_create_arraybuffer(env, n * sizeof(float), &data, &ab) != napi_ok) return NULL;
  if (n) memcpy(data, src, n * sizeof(float));
  if (napi_create_typedarray(env, napi_float32_array, n, ab, 0, &ta) != napi_ok) return NULL;
  return ta;
}
static napi_value make_f64(napi_env env, const double* src, size_t n) {
  void* data = NULL;
  napi_value ab, ta;
  if (napi_create_arraybuffer(env, n * sizeof(double), &data, &ab) != napi_ok) return NULL;
  if (n) memcpy(data, src, n * sizeof(double));
  if (napi_create_typedarray(env, napi_float64_array, n, ab, 0, &ta) != napi_ok) return NULL;
  return ta;
}
static void set_i(napi_env env, napi_value o, const char* k, int32_t v) { napi_value x; napi_create_int32(env, v, &x); napi_set_named_property(env, o, k, x); }
static void set_d(napi_env env, napi_value o, const char* k, double v) { napi_value x; napi_create_double(env, v, &x); napi_set_named_property(env, o, k, x); }

static void job_complete(napi_env env, napi_status status, void* data) {
  job_t* j = (job_t*)data;
  napi_value res;
  if (status != napi_ok || j->rc != FA_OK) {
    napi_create_string_utf8(env, j->err[0] ? j->err : "analysis failed", NAPI_AUTO_LENGTH, &res);
    napi_reject_deferred(env, j->deferred, res);
  } else {
    const fa_counts* c = &j->counts;
    napi_value counts, segs, syls;
    napi_create_object(env, &res);
    napi_create_object(env, &counts);
    set_i(env, counts, "frames", c->frames); set_i(env, counts, "bands", c->bands); set_i(env, counts, "hop", c->hop);
    set_i(env, counts, "sampleRate", c->sample_rate); set_i(env, counts, "segments", c->segments);
    set_i(env, counts, "syllables", c->syllables); set_i(env, counts, "featureRows", c->feature_rows);
    napi_set_named_property(env, res, "counts", counts);
    napi_create_array_with_length(env, (size_t)c->segments, &segs);
    for (int i = 0; i < c->segments; i++) {
      napi_value o;
      napi_create_object(env, &o);
      set_i(env, o, "start", j->segs[i].start); set_i(env, o, "len", j->segs[i].len); set_i(env, o, "stored", j->segs[i].stored);
      set_i(env, o, "nSyllables", j->segs[i].n_syllables); set_i(env, o, "firstSyllable", j->segs[i].first_syllable);
      set_i(env, o, "rowOffset", j->segs[i].row_offset); set_d(env, o, "ymax", j->segs[i].ymax);
      set_d(env, o, "vmin", j->segs[i].vmin); set_d(env, o, "csRatio", j->segs[i].cs_ratio);
      napi_set_element(env, segs, (uint32_t)i, o);
    }
    napi_set_named_property(env, res, "segments", segs);
    napi_create_array_with_length(env, (size_t)c->syllables, &syls);
    for (int i = 0; i < c->syllables; i++) {
      napi_value o;
      napi_create_object(env, &o);
      set_i(env, o, "storedSeg", j->syls[i].stored_seg); set_i(env, o, "start", j->syls[i].start); set_i(env, o, "len", j->syls[i].len);
      set_i(env, o, "flag", j->syls[i].reserved);     /* level 12: make_coeffs threw at or before this syllable */
      napi_set_element(env, syls, (uint32_t)i, o);
    }
    napi_set_named_property(env, res, "syllables", syls);     /* level 3: the fa_track headers (start = first point, len = points) */
    const size_t frows = j->level == FA_LEVEL_SEGMENTS ? 0 : (size_t)c->formant_rows;
    napi_set_named_property(env, res, "formants", make_f32(env, j->formants, 9 * frows));
    napi_set_named_property(env, res, "energy", make_f32(env, j->energy, 3 * frows));
    if (j->level == FA_LEVEL_SEGMENTS) {
      /* track points as six parallel columns: frame, lo, hi, bin, amp, energy */
      const size_t np = (size_t)c->formant_rows;
      double* col = (double*)malloc(sizeof(double) * 6 * (np + 1));
      if (col) {
        for (size_t i = 0; i < np; i++) {
          col[i] = j->points[i].frame; col[np + i] = j->points[i].lo; col[2 * np + i] = j->points[i].hi;
          col[3 * np + i] = j->points[i].bin; col[4 * np + i] = j->points[i].amp; col[5 * np + i] = j->points[i].energy;
        }
        napi_set_named_property(env, res, "trackPoints", make_f64(env, col, 6 * np));
        free(col);
      }
    }
    napi_set_named_property(env, res, "features", make_f64(env, j->features, (size_t)j->feat_width * (size_t)c->feature_rows));
    if (j->want_spec) napi_set_named_property(env, res, "spectrum", make_f32(env, j->spectrum, (size_t)j->fft_half * (size_t)c->frames));
    napi_resolve_deferred(env, j->deferred, res);
  }
  napi_delete_reference(env, j->pcm_ref);
  napi_delete_async_work(env, j->work);
  free(j->segs); free(j->syls); free(j->formants); free(j->energy); free(j->features); free(j->spectrum); free(j->points);
  free(j);
}

/* analyze(handle, Float32Array pcm, sampleRate, wantSpectrum, fftSize, outputLevel) -> Promise
 * analyze(handle, Uint32Array frames, bands, false, fftSize, outputLevel)       -> Promise (frames = spectrum_push input) */
static napi_value Analyze(napi_env env, napi_callback_info info) {
  size_t argc = 6;
  napi_value argv[6], promise, name;
  CHECK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 3) { napi_throw_error(env, NULL, "Invalid audio source"); return NULL; }
  job_t* j = (job_t*)calloc(1, sizeof(job_t));
  if (!j) { napi_throw_error(env, NULL, "out of memory"); return NULL; }
  void* hp = NULL;
  napi_typedarray_type tt;
  void* pdata = NULL;
  size_t len = 0;
  bool is_ta = false;
  if (napi_get_value_external(env, argv[0], &hp) != napi_ok || !hp || !((engine_box*)hp)->h || napi_is_typedarray(env, argv[1], &is_ta) != napi_ok || !is_ta ||
      napi_get_typedarray_info(env, argv[1], &tt, &len, &pdata, NULL, NULL) != napi_ok ||
      (tt != napi_float32_array && tt != napi_uint32_array)) {
    free(j);
    napi_throw_error(env, NULL, "Invalid audio source");
    return NULL;
  }
  int32_t sr = 0, ws = 0, fft = 2048;
  napi_get_value_int32(env, argv[2], &sr);
  if (argc >= 4) { bool b = false; if (napi_get_value_bool(env, argv[3], &b) == napi_ok) ws = b; }
  if (argc >= 5) napi_get_value_int32(env, argv[4], &fft);
  int32_t level = 0;
  if (argc >= 6) napi_get_value_int32(env, argv[5], &level);
  j->h = ((engine_box*)hp)->h; j->pcm = (const float*)pdata; j->n = len; j->sr = sr; j->want_spec = ws; j->fft_half = fft / 2;
  j->feat_width = level == FA_LEVEL_UTTERANCE ? FA_N_UTT_FEATURES : FA_N_FEATURES;
  j->is_frames = tt == napi_uint32_array;   /* Uint32Array = the segmentor's own frames (spectrum_push); argv[2] = bands */
  CHECK(napi_create_reference(env, argv[1], 1, &j->pcm_ref)); /* keep the PCM alive while the pool thread reads it */
  CHECK(napi_create_promise(env, &j->deferred, &promise));
  CHECK(napi_create_string_utf8(env, "fa_b200.analyze", NAPI_AUTO_LENGTH, &name));
  CHECK(napi_create_async_work(env, NULL, name, job_execute, job_complete, j, &j->work));
  CHECK(napi_queue_async_work(env, j->work));
  return promise;
}

#ifdef __cplusplus
extern "C"
#endif
__attribute__((visibility("default"))) napi_value napi_register_module_v1(napi_env env, napi_value exports) {
  napi_value f;
  if (napi_create_function(env, "createEngine", NAPI_AUTO_LENGTH, CreateEngine, NULL, &f) == napi_ok) napi_set_named_property(env, exports, "createEngine", f);
  if (napi_create_function(env, "destroyEngine", NAPI_AUTO_LENGTH, DestroyEngine, NULL, &f) == napi_ok) napi_set_named_property(env, exports, "destroyEngine", f);
  if (napi_create_function(env, "analyze", NAPI_AUTO_LENGTH, Analyze, NULL, &f) == napi_ok) napi_set_named_property(env, exports, "analyze", f);
  if (napi_create_function(env, "setTruncate", NAPI_AUTO_LENGTH, SetTruncate, NULL, &f) == napi_ok) napi_set_named_property(env, exports, "setTruncate", f);
  return exports;
}
