// esp_audio_b200.hpp — header-only C++ shim: the reference's names on top of the C ABI.
//
// The reference declares its hot-path API inside C++ namespaces (esp_audio_libs::art_resampler,
// ::quantization_utils, ::resampler; include/art_resampler.h:18-48, include/art_biquad.h:16-38,
// include/quantization_utils.h:6-28, include/resampler.h:10-82) and links it statically.
// This shim re-declares the same names in namespace esp_audio_libs_b200 with a batch
// dimension (num_streams) and a CUDA stream added, so that a caller switches by changing the
// namespace, passing device pointers and telling the context how many streams a call covers.
// Everything forwards to extern "C" entry points of libesp_audio_b200.so; nothing is computed here.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "esp_audio_b200.h"

namespace esp_audio_libs_b200 {

namespace art_resampler {

constexpr int SUBSAMPLE_INTERPOLATE_ = ESPB_SUBSAMPLE_INTERPOLATE;  // include/art_resampler.h:21-23
constexpr int BLACKMAN_HARRIS_ = ESPB_BLACKMAN_HARRIS;
constexpr int INCLUDE_LOWPASS_ = ESPB_INCLUDE_LOWPASS;

using Resample = EspbResampleBatch;          // opaque, like the reference's context pointer
using ResampleResult = EspbResampleResult;   // {input_used, output_generated}
using BiquadCoefficients = EspbBiquadCoefficients;
using Biquad = EspbBiquadBatch;              // a bank of Biquad states on the device

// include/art_resampler.h:35 — plus the number of streams the context stands for
inline Resample *resampleInit(int numStreams, int numChannels, int numTaps, int numFilters, float lowpassRatio,
                              int flags) {
  return espb_resampleInit(numStreams, numChannels, numTaps, numFilters, lowpassRatio, flags);
}
// include/art_resampler.h:36-39 — device pointers, row strides in floats, CUDA stream
inline ResampleResult resampleProcess(Resample *cxt, const float *input, int64_t inStreamStride,
                                      int64_t inChannelStride, int numInputFrames, float *output,
                                      int64_t outStreamStride, int64_t outChannelStride, int numOutputFrames,
                                      float ratio, void *stream = nullptr) {
  return espb_resampleProcess(cxt, input, inStreamStride, inChannelStride, numInputFrames, output, outStreamStride,
                              outChannelStride, numOutputFrames, ratio, stream);
}
// include/art_resampler.h:36-37 verbatim — one (device) pointer per plane, numStreams x numChannels of them
inline ResampleResult resampleProcess(Resample *cxt, const float *const *inputs, int numInputFrames,
                                      float *const *outputs, int numOutputFrames, float ratio,
                                      void *stream = nullptr) {
  return espb_resampleProcessPlanes(cxt, inputs, numInputFrames, outputs, numOutputFrames, ratio, stream);
}
inline ResampleResult resampleProcessInterleaved(Resample *cxt, const float *input, int64_t inStreamStride,
                                                 int numInputFrames, float *output, int64_t outStreamStride,
                                                 int numOutputFrames, float ratio, void *stream = nullptr) {
  return espb_resampleProcessInterleaved(cxt, input, inStreamStride, numInputFrames, output, outStreamStride,
                                         numOutputFrames, ratio, stream);
}
inline unsigned int resampleGetRequiredSamples(Resample *cxt, int numOutputFrames, float ratio) {
  return espb_resampleGetRequiredSamples(cxt, numOutputFrames, ratio);
}
inline unsigned int resampleGetExpectedOutput(Resample *cxt, int numInputFrames, float ratio) {
  return espb_resampleGetExpectedOutput(cxt, numInputFrames, ratio);
}
inline void resampleAdvancePosition(Resample *cxt, float delta) { espb_resampleAdvancePosition(cxt, delta); }
inline float resampleGetPosition(Resample *cxt) { return espb_resampleGetPosition(cxt); }
inline void resampleReset(Resample *cxt, void *stream = nullptr) { espb_resampleReset(cxt, stream); }
inline void resampleFree(Resample *cxt) { espb_resampleFree(cxt); }

// include/art_biquad.h:30-36
inline void biquad_lowpass(BiquadCoefficients *filter, double frequency) { espb_biquad_lowpass(filter, frequency); }
inline void biquad_highpass(BiquadCoefficients *filter, double frequency) { espb_biquad_highpass(filter, frequency); }
inline Biquad *biquad_init(int numSeries, int numSections, const BiquadCoefficients *coeffs, float gain) {
  return espb_biquad_init(numSeries, numSections, coeffs, gain);
}
// in place on interleaved device data: series q = (stream q / channels, channel q % channels)
inline int biquad_apply_buffer(Biquad *f, float *buffer, int64_t streamStride, int channels, int num_samples,
                               void *stream = nullptr) {
  EspbLayout l = {streamStride, 1, channels};
  return espb_biquad_apply_buffer(f, buffer, &l, channels, num_samples, stream);
}
// include/art_biquad.h:36 — one new sample per series of the bank (device memory, in place)
inline int biquad_apply_sample(Biquad *f, float *samples, void *stream = nullptr) {
  return espb_biquad_apply_samples(f, samples, stream);
}
inline void biquad_free(Biquad *f) { espb_biquad_free(f); }

}  // namespace art_resampler

namespace quantization_utils {

// include/quantization_utils.h:15-16
inline void quantized_to_float(const uint8_t *input_buffer, float *output_buffer, uint64_t num_samples,
                               uint8_t input_bits, float gain_db, void *stream = nullptr) {
  espb_quantized_to_float(input_buffer, output_buffer, num_samples, input_bits, gain_db, stream);
}
// include/quantization_utils.h:24-25 — returns the number of clipped samples (synchronises the stream)
inline uint32_t float_to_quantized(const float *input_buffer, uint8_t *output_buffer, uint64_t num_samples,
                                   uint8_t output_bits, void *stream = nullptr) {
  return espb_float_to_quantized_sync(input_buffer, output_buffer, num_samples, output_bits, stream);
}

}  // namespace quantization_utils

namespace resampler {

using ResamplerResults = EspbResamplerResults;              // include/resampler.h:15-20
using ResamplerConfiguration = EspbResamplerConfiguration;  // include/resampler.h:22-32

// include/resampler.h:36-80
class Resampler {
 public:
  Resampler(int num_streams, size_t input_buffer_samples, size_t output_buffer_samples)
      : num_streams_(num_streams),
        input_buffer_samples_(input_buffer_samples),
        output_buffer_samples_(output_buffer_samples) {}
  ~Resampler() { espb_resampler_free(impl_); }
  Resampler(const Resampler &) = delete;
  Resampler &operator=(const Resampler &) = delete;

  /// @return true if everything was allocated, false otherwise (as the reference)
  bool initialize(ResamplerConfiguration &config) {
    impl_ = espb_resampler_create(num_streams_, input_buffer_samples_, output_buffer_samples_, &config);
    return impl_ != nullptr;
  }
  /// device buffers: row s = stream s, rows *_stride_bytes apart
  ResamplerResults resample(const uint8_t *input_buffer, int64_t in_stride_bytes, uint8_t *output_buffer,
                            int64_t out_stride_bytes, size_t input_frames_available, size_t output_frames_free,
                            float gain_db, void *stream = nullptr) {
    return espb_resampler_resample(impl_, input_buffer, in_stride_bytes, output_buffer, out_stride_bytes,
                                   input_frames_available, output_frames_free, gain_db, nullptr, stream);
  }
  /// device buffers, enqueue only: frame counts are returned at once, clipped_samples is 0 (the per-stream counts
  /// are at clipped_dev() once the stream has reached this point)
  ResamplerResults resample_async(const uint8_t *input_buffer, int64_t in_stride_bytes, uint8_t *output_buffer,
                                  int64_t out_stride_bytes, size_t input_frames_available, size_t output_frames_free,
                                  float gain_db, void *stream = nullptr) {
    return espb_resampler_resample_async(impl_, input_buffer, in_stride_bytes, output_buffer, out_stride_bytes,
                                         input_frames_available, output_frames_free, gain_db, stream);
  }
  const uint32_t *clipped_dev() { return espb_resampler_clipped_dev(impl_); }
  /// host buffers (staged, copied and pipelined internally)
  ResamplerResults resample_host(const uint8_t *input_buffer, int64_t in_stride_bytes, uint8_t *output_buffer,
                                 int64_t out_stride_bytes, size_t input_frames_available,
                                 size_t output_frames_free, float gain_db) {
    return espb_resampler_resample_host(impl_, input_buffer, in_stride_bytes, output_buffer, out_stride_bytes,
                                        input_frames_available, output_frames_free, gain_db, nullptr);
  }

 protected:
  int num_streams_;
  size_t input_buffer_samples_, output_buffer_samples_;
  EspbResampler *impl_{nullptr};
};

}  // namespace resampler
namespace wav_decoder {

// include/wav_decoder.h:34-52 — same enumerators, same values
enum WAVDecoderState {
  WAV_DECODER_BEFORE_RIFF = ESPB_WAV_DECODER_BEFORE_RIFF,
  WAV_DECODER_BEFORE_WAVE = ESPB_WAV_DECODER_BEFORE_WAVE,
  WAV_DECODER_BEFORE_FMT = ESPB_WAV_DECODER_BEFORE_FMT,
  WAV_DECODER_IN_FMT = ESPB_WAV_DECODER_IN_FMT,
  WAV_DECODER_BEFORE_DATA = ESPB_WAV_DECODER_BEFORE_DATA,
  WAV_DECODER_IN_DATA = ESPB_WAV_DECODER_IN_DATA,
};
enum WAVDecoderResult {
  WAV_DECODER_SUCCESS_NEXT = ESPB_WAV_DECODER_SUCCESS_NEXT,
  WAV_DECODER_SUCCESS_IN_DATA = ESPB_WAV_DECODER_SUCCESS_IN_DATA,
  WAV_DECODER_WARNING_INCOMPLETE_DATA = ESPB_WAV_DECODER_WARNING_INCOMPLETE_DATA,
  WAV_DECODER_ERROR_NO_RIFF = ESPB_WAV_DECODER_ERROR_NO_RIFF,
  WAV_DECODER_ERROR_NO_WAVE = ESPB_WAV_DECODER_ERROR_NO_WAVE,
  WAV_DECODER_ERROR_FAILED = ESPB_WAV_DECODER_ERROR_FAILED,
};

// include/wav_decoder.h:54-89 (host code; no device involved)
class WAVDecoder {
 public:
  WAVDecoder() : impl_(espb_wav_decoder_create()) {}
  ~WAVDecoder() { espb_wav_decoder_free(impl_); }
  WAVDecoder(const WAVDecoder &) = delete;
  WAVDecoder &operator=(const WAVDecoder &) = delete;

  WAVDecoderState state() { return (WAVDecoderState) espb_wav_decoder_state(impl_); }
  std::size_t bytes_processed() { return espb_wav_decoder_bytes_processed(impl_); }
  std::size_t bytes_to_skip() { return espb_wav_decoder_bytes_to_skip(impl_); }
  std::size_t bytes_needed() { return espb_wav_decoder_bytes_needed(impl_); }
  std::string chunk_name() { return std::string(espb_wav_decoder_chunk_name(impl_), 4); }
  std::size_t chunk_bytes_left() { return espb_wav_decoder_chunk_bytes_left(impl_); }
  uint32_t sample_rate() { return espb_wav_decoder_sample_rate(impl_); }
  uint16_t num_channels() { return espb_wav_decoder_num_channels(impl_); }
  uint16_t bits_per_sample() { return espb_wav_decoder_bits_per_sample(impl_); }

  WAVDecoderResult decode_header(const uint8_t *buffer, size_t bytes_available) {
    return (WAVDecoderResult) espb_wav_decoder_decode_header(impl_, buffer, bytes_available);
  }
  WAVDecoderResult next(const uint8_t *buffer) { return (WAVDecoderResult) espb_wav_decoder_next(impl_, buffer); }
  void reset() { espb_wav_decoder_reset(impl_); }

 protected:
  EspbWavDecoder *impl_;
};

}  // namespace wav_decoder

// include/dsp.h:66-93,109-115 — the portable-C Q15 helpers on device buffers; 0 = ESP_OK, -1 = ESP_FAIL
inline int dsps_add_s16(const int16_t *input1, const int16_t *input2, int16_t *output, int64_t len, int step1,
                        int step2, int step_out, int shift, void *stream = nullptr) {
  return espb_dsps_add_s16(input1, input2, output, len, step1, step2, step_out, shift, stream) == ESPB_OK ? 0 : -1;
}
inline int dsps_mulc_s16(const int16_t *input, int16_t *output, int64_t len, int16_t C, int step_in, int step_out,
                         void *stream = nullptr) {
  return espb_dsps_mulc_s16(input, output, len, C, step_in, step_out, stream) == ESPB_OK ? 0 : -1;
}

}  // namespace esp_audio_libs_b200
