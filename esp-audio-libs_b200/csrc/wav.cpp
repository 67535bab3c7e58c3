// WAV header parse / emit: the on-disk format either side of the resampler path.  Host code
// only.  The parser is the reference's incremental state machine (include/wav_decoder.h:32-90,
// src/decode/wav_decoder.cpp:8-161) with the same states, result codes, byte accounting and quirks:
//   * chunk sizes are rounded up to even (RIFF pad byte) before skipping/reading (:61-65, :85-89, :125-129);
//   * the fmt chunk is read whole, whatever its size, and fields are taken at offsets 2, 4 and 14 (:111-113);
//   * reset() (:151-161) restores the state and the parsed fields but NOT bytes_needed — a decoder reset after it
//     reached the data chunk reports WAV_DECODER_ERROR_FAILED on the next decode_header, as the reference does;
//     espb_wav_decoder_create() after espb_wav_decoder_free() is the way to start over.
// The writer emits the canonical 44-byte PCM header those parsers accept (the reference has no writer).
#include <stdint.h>
#include <stdlib.h>

#include <new>
#include <string.h>

#include "../../include/esp_audio_b200.h"

struct EspbWavDecoder {
  int state = ESPB_WAV_DECODER_BEFORE_RIFF;
  size_t bytes_processed = 0;
  size_t bytes_needed = 8;  // chunk name + size
  size_t bytes_to_skip = 0;
  size_t chunk_bytes_left = 0;
  char chunk_name[5] = {0, 0, 0, 0, 0};
  uint32_t sample_rate = 0;
  uint16_t num_channels = 0, bits_per_sample = 0;
};

namespace {

inline bool is_tag(const EspbWavDecoder *d, const char *tag) { return memcmp(d->chunk_name, tag, 4) == 0; }

// the 32-bit little-endian size lands in the low half of the size_t member, then the pad byte
inline void take_chunk_size(EspbWavDecoder *d, const uint8_t *p) {
  const uint32_t v = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24);
  d->chunk_bytes_left = (d->chunk_bytes_left & ~(size_t) 0xffffffffu) | v;
  if (d->chunk_bytes_left % 2 != 0)
    d->chunk_bytes_left++;
}

inline void put16(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t) v;
  p[1] = (uint8_t) (v >> 8);
}
inline void put32(uint8_t *p, uint32_t v) {
  put16(p, v);
  put16(p + 2, v >> 16);
}

}  // namespace

extern "C" {

EspbWavDecoder *espb_wav_decoder_create(void) { return new (std::nothrow) EspbWavDecoder(); }
void espb_wav_decoder_free(EspbWavDecoder *d) { delete d; }

int espb_wav_decoder_next(EspbWavDecoder *d, const uint8_t *buffer) {
  d->bytes_to_skip = 0;
  switch (d->state) {
    case ESPB_WAV_DECODER_BEFORE_RIFF:
      memcpy(d->chunk_name, buffer, 4);
      if (!is_tag(d, "RIFF"))
        return ESPB_WAV_DECODER_ERROR_NO_RIFF;
      take_chunk_size(d, buffer + 4);
      d->state = ESPB_WAV_DECODER_BEFORE_WAVE;
      d->bytes_needed = 4;
      return ESPB_WAV_DECODER_SUCCESS_NEXT;
    case ESPB_WAV_DECODER_BEFORE_WAVE:
      memcpy(d->chunk_name, buffer, 4);
      if (!is_tag(d, "WAVE"))
        return ESPB_WAV_DECODER_ERROR_NO_WAVE;
      d->state = ESPB_WAV_DECODER_BEFORE_FMT;
      d->bytes_needed = 8;
      return ESPB_WAV_DECODER_SUCCESS_NEXT;
    case ESPB_WAV_DECODER_BEFORE_FMT:
    case ESPB_WAV_DECODER_BEFORE_DATA: {
      const bool want_fmt = d->state == ESPB_WAV_DECODER_BEFORE_FMT;
      memcpy(d->chunk_name, buffer, 4);
      take_chunk_size(d, buffer + 4);
      if (want_fmt && is_tag(d, "fmt ")) {
        d->state = ESPB_WAV_DECODER_IN_FMT;
        d->bytes_needed = d->chunk_bytes_left;
      } else if (!want_fmt && is_tag(d, "data")) {
        d->state = ESPB_WAV_DECODER_IN_DATA;
        d->bytes_needed = 0;
        return ESPB_WAV_DECODER_SUCCESS_IN_DATA;
      } else {  // LIST, INFO, ...: skip the whole chunk, then expect another chunk header
        d->bytes_to_skip = d->chunk_bytes_left;
        d->bytes_needed = 8;
      }
      return ESPB_WAV_DECODER_SUCCESS_NEXT;
    }
    case ESPB_WAV_DECODER_IN_FMT:
      d->num_channels = (uint16_t) (buffer[2] | (buffer[3] << 8));
      d->sample_rate = (uint32_t) buffer[4] | ((uint32_t) buffer[5] << 8) | ((uint32_t) buffer[6] << 16) |
                       ((uint32_t) buffer[7] << 24);
      d->bits_per_sample = (uint16_t) (buffer[14] | (buffer[15] << 8));
      d->state = ESPB_WAV_DECODER_BEFORE_DATA;
      d->bytes_needed = 8;
      return ESPB_WAV_DECODER_SUCCESS_NEXT;
    default:
      return ESPB_WAV_DECODER_SUCCESS_IN_DATA;
  }
}

int espb_wav_decoder_decode_header(EspbWavDecoder *d, const uint8_t *buffer, size_t bytes_available) {
  size_t to_skip = d->bytes_to_skip, to_read = d->bytes_needed;
  d->bytes_processed = 0;
  while (to_skip + to_read > 0) {
    if (to_skip > bytes_available || to_read > bytes_available)
      return ESPB_WAV_DECODER_WARNING_INCOMPLETE_DATA;
    if (to_skip > 0) {
      buffer += to_skip;
      d->bytes_processed += to_skip;
      bytes_available -= to_skip;
      to_skip = 0;
      continue;
    }
    const int result = espb_wav_decoder_next(d, buffer);
    buffer += to_read;
    d->bytes_processed += to_read;
    bytes_available -= to_read;
    if (result != ESPB_WAV_DECODER_SUCCESS_NEXT)
      return result;  // in the data chunk, or a malformed header
    to_skip = d->bytes_to_skip;
    to_read = d->bytes_needed;
  }
  return ESPB_WAV_DECODER_ERROR_FAILED;
}

void espb_wav_decoder_reset(EspbWavDecoder *d) {
  d->state = ESPB_WAV_DECODER_BEFORE_RIFF;
  d->bytes_to_skip = 0;
  memset(d->chunk_name, 0, sizeof d->chunk_name);
  d->chunk_bytes_left = 0;
  d->sample_rate = 0;
  d->num_channels = 0;
  d->bits_per_sample = 0;
}

int espb_wav_decoder_state(const EspbWavDecoder *d) { return d->state; }
size_t espb_wav_decoder_bytes_processed(const EspbWavDecoder *d) { return d->bytes_processed; }
size_t espb_wav_decoder_bytes_to_skip(const EspbWavDecoder *d) { return d->bytes_to_skip; }
size_t espb_wav_decoder_bytes_needed(const EspbWavDecoder *d) { return d->bytes_needed; }
const char *espb_wav_decoder_chunk_name(const EspbWavDecoder *d) { return d->chunk_name; }
size_t espb_wav_decoder_chunk_bytes_left(const EspbWavDecoder *d) { return d->chunk_bytes_left; }
uint32_t espb_wav_decoder_sample_rate(const EspbWavDecoder *d) { return d->sample_rate; }
uint16_t espb_wav_decoder_num_channels(const EspbWavDecoder *d) { return d->num_channels; }
uint16_t espb_wav_decoder_bits_per_sample(const EspbWavDecoder *d) { return d->bits_per_sample; }

size_t espb_wav_write_header(uint8_t *dst, uint32_t sample_rate, uint16_t num_channels, uint16_t bits_per_sample,
                             uint32_t data_bytes) {
  if (!dst)
    return 0;
  const uint32_t block_align = (uint32_t) num_channels * ((bits_per_sample + 7u) / 8u);
  const uint32_t padded = data_bytes + (data_bytes & 1u);
  memcpy(dst, "RIFF", 4);
  put32(dst + 4, 36u + padded);
  memcpy(dst + 8, "WAVE", 4);
  memcpy(dst + 12, "fmt ", 4);
  put32(dst + 16, 16);
  put16(dst + 20, 1);  // PCM
  put16(dst + 22, num_channels);
  put32(dst + 24, sample_rate);
  put32(dst + 28, sample_rate * block_align);
  put16(dst + 32, block_align);
  put16(dst + 34, bits_per_sample);
  memcpy(dst + 36, "data", 4);
  put32(dst + 40, data_bytes);
  return 44;
}

}  // extern "C"
