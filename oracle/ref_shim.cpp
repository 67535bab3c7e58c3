// TEST INFRASTRUCTURE — not part of the product.
//
// extern "C" shim over the UNMODIFIED reference sources under /root/reference
// (compiled in place by oracle/Makefile into oracle/_ref/libesp_audio_ref.so).
// It only forwards to the reference's own public functions so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// can call the real reference through ctypes.  Nothing under the product
// package may load this library.
//
// Forwarded interfaces (reference file:line):
//   include/art_resampler.h:35-45   resampleInit ... resampleFree
//   include/art_biquad.h:30-36      biquad_init / lowpass / highpass / apply
//   include/quantization_utils.h:15-25
//   include/resampler.h:36-80       resampler::Resampler
//   include/dsp.h:66-93             dsps_add_s16_ansi / dsps_mulc_s16_ansi
//   include/wav_decoder.h:54-89     wav_decoder::WAVDecoder
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

#include <vector>

#include "art_biquad.h"
#include "art_resampler.h"
#include "dsp.h"
#include "quantization_utils.h"
#include "resampler.h"
#include "wav_decoder.h"

using namespace esp_audio_libs;
namespace art = esp_audio_libs::art_resampler;

extern "C" {

// ---- art_resampler -------------------------------------------------------
void *ref_resampleInit(int ch, int taps, int filters, float lowpass, int flags) {
  return art::resampleInit(ch, taps, filters, lowpass, flags);
}
void ref_resampleFree(void *c) { art::resampleFree((art::Resample *) c); }
void ref_resampleReset(void *c) { art::resampleReset((art::Resample *) c); }
void ref_resampleAdvancePosition(void *c, float d) { art::resampleAdvancePosition((art::Resample *) c, d); }
float ref_resampleGetPosition(void *c) { return art::resampleGetPosition((art::Resample *) c); }
unsigned ref_resampleGetRequiredSamples(void *c, int n_out, float ratio) {
  return art::resampleGetRequiredSamples((art::Resample *) c, n_out, ratio);
}
unsigned ref_resampleGetExpectedOutput(void *c, int n_in, float ratio) {
  return art::resampleGetExpectedOutput((art::Resample *) c, n_in, ratio);
}
void ref_resampleProcessInterleaved(void *c, const float *in, int n_in, float *out, int n_out, float ratio,
                                    unsigned *used, unsigned *generated) {
  art::ResampleResult r = art::resampleProcessInterleaved((art::Resample *) c, in, n_in, out, n_out, ratio);
  *used = r.input_used;
  *generated = r.output_generated;
}
void ref_resampleProcess(void *c, const float *const *in, int n_in, float *const *out, int n_out, float ratio,
                         unsigned *used, unsigned *generated) {
  art::ResampleResult r = art::resampleProcess((art::Resample *) c, in, n_in, out, n_out, ratio);
  *used = r.input_used;
  *generated = r.output_generated;
}
// Introspection of the context (plain reads of the public struct).
int ref_resampleNumFilters(void *c) { return ((art::Resample *) c)->numFilters; }
int ref_resampleFlags(void *c) { return ((art::Resample *) c)->flags; }
void ref_resampleCopyFilter(void *c, int idx, float *dst) {
  art::Resample *r = (art::Resample *) c;
  memcpy(dst, r->filters[idx], r->numTaps * sizeof(float));
}
void ref_resampleState(void *c, float *output_offset, int *input_index) {
  *output_offset = ((art::Resample *) c)->outputOffset;
  *input_index = ((art::Resample *) c)->inputIndex;
}

// ---- art_biquad ------------------------------------------------------------
// coeffs: float[5] = a0 a1 a2 b1 b2 ; state: the reference's Biquad struct (opaque, 40 bytes)
int ref_biquad_sizeof(void) { return (int) sizeof(art::Biquad); }
void ref_biquad_lowpass(float *coeffs, double f) { art::biquad_lowpass((art::BiquadCoefficients *) coeffs, f); }
void ref_biquad_highpass(float *coeffs, double f) { art::biquad_highpass((art::BiquadCoefficients *) coeffs, f); }
void ref_biquad_init(void *bq, const float *coeffs, float gain) {
  art::biquad_init((art::Biquad *) bq, (const art::BiquadCoefficients *) coeffs, gain);
}
void ref_biquad_apply_buffer(void *bq, float *buf, int n, int stride) {
  art::biquad_apply_buffer((art::Biquad *) bq, buf, n, stride);
}
float ref_biquad_apply_sample(void *bq, float x) { return art::biquad_apply_sample((art::Biquad *) bq, x); }

// ---- quantization_utils ----------------------------------------------------
void ref_quantized_to_float(const uint8_t *in, float *out, uint32_t n, uint8_t bits, float gain_db) {
  quantization_utils::quantized_to_float(in, out, n, bits, gain_db);
}
uint32_t ref_float_to_quantized(const float *in, uint8_t *out, uint32_t n, uint8_t bits) {
  return quantization_utils::float_to_quantized(in, out, n, bits);
}

// ---- dsp.h Q15 helpers (portable C versions) ---------------------------------
int ref_add_s16(const int16_t *in1, const int16_t *in2, int16_t *out, int len, int step1, int step2, int step_out,
                int shift) {
  return dsps_add_s16_ansi(in1, in2, out, len, step1, step2, step_out, shift);
}
int ref_mulc_s16(const int16_t *in, int16_t *out, int len, int16_t c, int step_in, int step_out) {
  return dsps_mulc_s16_ansi(in, out, len, c, step_in, step_out);
}

// ---- wav_decoder::WAVDecoder --------------------------------------------------
void *ref_wav_create(void) { return new esp_audio_libs::wav_decoder::WAVDecoder(); }
void ref_wav_free(void *w) { delete (esp_audio_libs::wav_decoder::WAVDecoder *) w; }
int ref_wav_next(void *w, const uint8_t *buffer) { return (int) ((esp_audio_libs::wav_decoder::WAVDecoder *) w)->next(buffer); }
int ref_wav_decode_header(void *w, const uint8_t *buffer, size_t n) {
  return (int) ((esp_audio_libs::wav_decoder::WAVDecoder *) w)->decode_header(buffer, n);
}
void ref_wav_reset(void *w) { ((esp_audio_libs::wav_decoder::WAVDecoder *) w)->reset(); }
void ref_wav_snapshot(void *p, uint64_t out[8], char name[5]) {
  esp_audio_libs::wav_decoder::WAVDecoder *w = (esp_audio_libs::wav_decoder::WAVDecoder *) p;
  out[0] = (uint64_t) w->state();
  out[1] = w->bytes_processed();
  out[2] = w->bytes_needed();
  out[3] = w->bytes_to_skip();
  out[4] = w->chunk_bytes_left();
  out[5] = w->sample_rate();
  out[6] = w->num_channels();
  out[7] = w->bits_per_sample();
  std::string n = w->chunk_name();
  memset(name, 0, 5);
  memcpy(name, n.data(), n.size() < 4 ? n.size() : 4);
}

// ---- resampler::Resampler wrapper -----------------------------------------
struct RefWrapper {
  resampler::Resampler *r;
};
void *ref_wrapper_create(size_t in_samples, size_t out_samples, float src_rate, float dst_rate, int src_bits,
                         int dst_bits, int channels, int use_filter, int interpolate, int taps, int filters) {
  resampler::Resampler *r = new resampler::Resampler(in_samples, out_samples);
  resampler::ResamplerConfiguration cfg;
  cfg.source_sample_rate = src_rate;
  cfg.target_sample_rate = dst_rate;
  cfg.source_bits_per_sample = (uint8_t) src_bits;
  cfg.target_bits_per_sample = (uint8_t) dst_bits;
  cfg.channels = (uint8_t) channels;
  cfg.use_pre_or_post_filter = use_filter != 0;
  cfg.subsample_interpolate = interpolate != 0;
  cfg.number_of_taps = (uint16_t) taps;
  cfg.number_of_filters = (uint16_t) filters;
  if (!r->initialize(cfg)) {
    delete r;
    return NULL;
  }
  return r;
}
void ref_wrapper_free(void *w) { delete (resampler::Resampler *) w; }
// results: uint64[4] = frames_used, frames_generated, predicted_frames_used, clipped_samples
void ref_wrapper_resample(void *w, const uint8_t *in, uint8_t *out, size_t in_frames, size_t out_free,
                          float gain_db, uint64_t *results) {
  resampler::ResamplerResults r = ((resampler::Resampler *) w)->resample(in, out, in_frames, out_free, gain_db);
  results[0] = r.frames_used;
  results[1] = r.frames_generated;
  results[2] = r.predicted_frames_used;
  results[3] = r.clipped_samples;
}

// ---- CPU baseline driver ---------------------------------------------------
// Runs resampleProcessInterleaved over `n_streams` independent contexts dealt
// round-robin to `n_threads` host threads; every stream reads its own input
// row (in + s*in_stride floats) and writes its own output row.  Contexts are
// created (and advanced by taps/2, as resampler.cpp:94 does) outside the timed
// region.  Returns seconds of wall-clock spent in processing only.
struct BenchJob {
  int first, step, n_streams;
  std::vector<void *> *ctx;
  const float *in;
  size_t in_stride;
  int n_in;
  float *out;
  size_t out_stride;
  int n_out;
  float ratio;
  unsigned long long generated;
};
static void *bench_thread(void *p) {
  BenchJob *j = (BenchJob *) p;
  unsigned long long g = 0;
  for (int s = j->first; s < j->n_streams; s += j->step) {
    art::ResampleResult r = art::resampleProcessInterleaved((art::Resample *) (*j->ctx)[s], j->in + s * j->in_stride,
                                                            j->n_in, j->out + s * j->out_stride, j->n_out, j->ratio);
    g += r.output_generated;
  }
  j->generated = g;
  return NULL;
}
double ref_bench_resample(int n_streams, int n_threads, int channels, int taps, int filters, float lowpass, int flags,
                          float advance, const float *in, size_t in_stride, int n_in, float *out, size_t out_stride,
                          int n_out, float ratio, unsigned long long *frames_generated) {
  std::vector<void *> ctx(n_streams);
  for (int s = 0; s < n_streams; ++s) {
    ctx[s] = art::resampleInit(channels, taps, filters, lowpass, flags);
    if (!ctx[s])
      return -1.0;
    if (advance > 0.0f)
      art::resampleAdvancePosition((art::Resample *) ctx[s], advance);
  }
  if (n_threads < 1)
    n_threads = 1;
  std::vector<BenchJob> jobs(n_threads);
  std::vector<pthread_t> tids(n_threads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < n_threads; ++t) {
    BenchJob &j = jobs[t];
    j.first = t;
    j.step = n_threads;
    j.n_streams = n_streams;
    j.ctx = &ctx;
    j.in = in;
    j.in_stride = in_stride;
    j.n_in = n_in;
    j.out = out;
    j.out_stride = out_stride;
    j.n_out = n_out;
    j.ratio = ratio;
    j.generated = 0;
    pthread_create(&tids[t], NULL, bench_thread, &j);
  }
  unsigned long long total = 0;
  for (int t = 0; t < n_threads; ++t) {
    pthread_join(tids[t], NULL);
    total += jobs[t].generated;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (int s = 0; s < n_streams; ++s)
    art::resampleFree((art::Resample *) ctx[s]);
  if (frames_generated)
    *frames_generated = total;
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

}  // extern "C"
