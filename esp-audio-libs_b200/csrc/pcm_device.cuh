// float -> packed little-endian PCM on the device (quantization_utils.cpp:50-94): the per-sample arithmetic and the
// packing of four samples into 32-bit words, shared by the quantiser kernels (pcm_kernels.cu) and the fused
// post-filter + quantiser kernel (biquad_kernel.cu).  Internal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace espb {
namespace {

struct F2QConst {
  float scalar;
  int32_t offset, hi, lo;
  int shift, bits;
};

__device__ __forceinline__ int32_t quantise_one(float x, const F2QConst &c, uint32_t &clipped) {
  int32_t v = __float2int_rd(__fadd_rn(__fmul_rn(x, c.scalar), 0.5f));  // :61 floorf(x*scalar + 0.5f)
  if (c.bits < 32) {                                                    // :62-69
    if (v > c.hi) {
      ++clipped;
      v = c.hi;
    } else if (v < c.lo) {
      ++clipped;
      v = c.lo;
    }
  } else {  // :70-78
    if (x >= 1.0f) {
      ++clipped;
      v = c.hi;
    } else if (x < -1.0f) {
      ++clipped;
      v = c.lo;
    }
  }
  return (int32_t) ((uint32_t) v << c.shift) + c.offset;  // :80
}

// The same arithmetic without branches (compare / select only) and with the 32-bit case chosen at compile time: for
// kernels in which the quantisation rides beside a latency-bound dependency chain (the fused post-filter), where a
// divergent branch per sample would keep the in-order warp from overlapping it with the next sample's recurrence.
template <bool BITS32>
__device__ __forceinline__ int32_t quantise_one_nb(float x, const F2QConst &c, uint32_t &clipped) {
  int32_t v = __float2int_rd(__fadd_rn(__fmul_rn(x, c.scalar), 0.5f));  // :61
  if (!BITS32) {                                                        // :62-69
    clipped += (uint32_t) (v > c.hi) + (uint32_t) (v < c.lo);
    v = min(max(v, c.lo), c.hi);
  } else {  // :70-78
    const bool over = x >= 1.0f, under = x < -1.0f;
    clipped += (uint32_t) over + (uint32_t) under;
    v = over ? c.hi : (under ? c.lo : v);
  }
  return (int32_t) ((uint32_t) v << c.shift) + c.offset;  // :80
}

template <int NBYTES>
__device__ __forceinline__ void encode_words(const int32_t v[4], uint32_t *w) {
  if (NBYTES == 1) {
    w[0] = ((uint32_t) v[0] & 0xffu) | (((uint32_t) v[1] & 0xffu) << 8) | (((uint32_t) v[2] & 0xffu) << 16) |
           ((uint32_t) v[3] << 24);
  } else if (NBYTES == 2) {
    w[0] = ((uint32_t) v[0] & 0xffffu) | ((uint32_t) v[1] << 16);
    w[1] = ((uint32_t) v[2] & 0xffffu) | ((uint32_t) v[3] << 16);
  } else if (NBYTES == 3) {
    const uint32_t a = (uint32_t) v[0] & 0xffffffu, b = (uint32_t) v[1] & 0xffffffu,
                   c = (uint32_t) v[2] & 0xffffffu, d = (uint32_t) v[3] & 0xffffffu;
    w[0] = a | (b << 24);
    w[1] = (b >> 8) | (c << 16);
    w[2] = (c >> 16) | (d << 8);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = (uint32_t) v[i];
  }
}

inline F2QConst make_f2q_const(int bits) {
  F2QConst c;
  c.bits = bits;
  c.scalar = (float) ((uint64_t) 1 << bits) / 2.0f;  // :52
  c.offset = (bits <= 8) ? 128 : 0;                   // :53
  c.hi = (int32_t) ((1u << (bits - 1)) - 1u);         // :54
  c.lo = ~c.hi;                                       // :55
  c.shift = (32 - bits) % 8;                          // :56
  return c;
}

}  // namespace
}  // namespace espb
