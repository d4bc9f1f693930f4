/*
 * fa_synth.h -- synthetic "glottal pulse through formant resonators" speech: the workload generator of
 * bench.py and the tests.  Host code in its own library (webspeechanalyzer_b200/libfa_synth.so), deterministic in
 * (seed, utt_index); it has no counterpart in the reference (which ships one demo WAV) and is not part of the drop-in boundary.
 */
#ifndef FA_SYNTH_H_
#define FA_SYNTH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FA_SYNTH_API __attribute__((visibility("default")))
#else
#define FA_SYNTH_API
#endif

/* n_samples of mono float32 speech at sample_rate; returns 0, or -1 on bad arguments */
FA_SYNTH_API int fa_synth_speech(float* dst, size_t n_samples, int sample_rate, uint64_t seed, uint64_t utt_index);
/* n_utt utterances of n_samples each as int16 PCM, back to back; utterance i has index first_index + i * index_stride */
FA_SYNTH_API int fa_synth_speech_i16_batch(int16_t* dst, int n_utt, size_t n_samples, int sample_rate, uint64_t seed,
                                           uint64_t first_index, uint64_t index_stride, int threads);

#ifdef __cplusplus
}
#endif
#endif /* FA_SYNTH_H_ */
