// art_biquad on the device: Direct-Form-I sections, sequential in time, one thread per
// series (stream x channel), thousands of series in flight.
//
// Reference: src/resample/art_biquad.cpp:73-93 (biquad_apply_buffer) — the sum
//   x*a0 + in_d1*a1 + in_d2*a2 - b1*out_d1 - b2*out_d2
// is evaluated left to right with every product and sum rounded (no FMA): contraction
// moves the result by up to 1.8e-6, so only __fmul_rn/__fadd_rn/
// __fsub_rn are used.  Cascaded sections are applied per sample instead of as separate
// buffer passes; the arithmetic per section is unchanged, so the result is bit-identical.
//
// Data movement: a CTA owns 32 consecutive series and walks time in tiles of 32 frames
// staged through shared memory, so that HBM sees the reference layout (stream-major,
// channels interleaved) as coalesced row segments while each thread reads its own series.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.hpp"

namespace espb {

namespace {

constexpr int kMaxSections = 4;
constexpr int BQ_SERIES = 32;  // series per CTA (one warp does the recurrences)
constexpr int BQ_TILE = 64;    // frames per tile
constexpr int BQ_THREADS = 128;

struct Section {
  float in_d1, in_d2, out_d1, out_d2;
};

__device__ __forceinline__ float section_step(Section &s, float x, const BiquadParams &c) {
  float sum = __fadd_rn(__fmul_rn(x, c.a0), __fmul_rn(s.in_d1, c.a1));
  if (!c.first_order)
    sum = __fadd_rn(sum, __fmul_rn(s.in_d2, c.a2));
  sum = __fsub_rn(sum, __fmul_rn(c.b1, s.out_d1));
  if (!c.first_order)
    sum = __fsub_rn(sum, __fmul_rn(c.b2, s.out_d2));
  s.out_d2 = s.out_d1;
  s.out_d1 = sum;
  s.in_d2 = s.in_d1;
  s.in_d1 = x;
  return sum;
}

template <int NSEC>
__global__ void __launch_bounds__(BQ_THREADS)
    espb_biquad_kernel(float *__restrict__ buf, int64_t ss, int64_t cs, int64_t fs, int channels, int n_series,
                       int n_samples, BiquadParams c, float *__restrict__ state) {
  __shared__ float tile[BQ_TILE][BQ_SERIES + 1];
  const int q0 = blockIdx.x * BQ_SERIES;
  const int tid = threadIdx.x;

  // loader mapping: element i of a tile -> (series, frame), chosen so that consecutive
  // threads touch consecutive addresses in the common layouts.
  // 0: frames contiguous (planar)  1: interleaved, CTA covers whole streams  2: generic
  const int mapping = (fs == 1) ? 0 : ((cs == 1 && fs == channels && BQ_SERIES % channels == 0) ? 1 : 2);
  auto locate = [&](int i, int &sl, int &t) {
    if (mapping == 0) {
      sl = i / BQ_TILE;
      t = i - sl * BQ_TILE;
    } else if (mapping == 1) {
      const int per = BQ_TILE * channels;
      const int stl = i / per, r = i - stl * per;
      t = r / channels;
      sl = stl * channels + (r - t * channels);
    } else {
      t = i / BQ_SERIES;
      sl = i - t * BQ_SERIES;
    }
  };

  Section sec[NSEC];
  const int q = q0 + tid;  // recurrence owner (threads 0..31)
  if (tid < BQ_SERIES && q < n_series) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
      const float4 v = *reinterpret_cast<const float4 *>(state + ((int64_t) q * NSEC + k) * 4);
      sec[k].in_d1 = v.x;
      sec[k].in_d2 = v.y;
      sec[k].out_d1 = v.z;
      sec[k].out_d2 = v.w;
    }
  }

  for (int t0 = 0; t0 < n_samples; t0 += BQ_TILE) {
    const int nt = (n_samples - t0 < BQ_TILE) ? n_samples - t0 : BQ_TILE;
    // ---- load tile
    for (int i = tid; i < BQ_SERIES * BQ_TILE; i += BQ_THREADS) {
      int sl, t;
      locate(i, sl, t);
      const int qq = q0 + sl;
      if (qq < n_series && t < nt) {
        const int st = qq / channels, ch = qq - st * channels;
        tile[t][sl] = buf[(int64_t) st * ss + (int64_t) ch * cs + (int64_t) (t0 + t) * fs];
      }
    }
    __syncthreads();
    // ---- recurrences
    if (tid < BQ_SERIES && q < n_series) {
      for (int t = 0; t < nt; ++t) {
        float v = tile[t][tid];
#pragma unroll
        for (int k = 0; k < NSEC; ++k)
          v = section_step(sec[k], v, c);
        tile[t][tid] = v;
      }
    }
    __syncthreads();
    // ---- store tile
    for (int i = tid; i < BQ_SERIES * BQ_TILE; i += BQ_THREADS) {
      int sl, t;
      locate(i, sl, t);
      const int qq = q0 + sl;
      if (qq < n_series && t < nt) {
        const int st = qq / channels, ch = qq - st * channels;
        buf[(int64_t) st * ss + (int64_t) ch * cs + (int64_t) (t0 + t) * fs] = tile[t][sl];
      }
    }
    __syncthreads();
  }

  if (tid < BQ_SERIES && q < n_series) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k)
      *reinterpret_cast<float4 *>(state + ((int64_t) q * NSEC + k) * 4) =
          make_float4(sec[k].in_d1, sec[k].in_d2, sec[k].out_d1, sec[k].out_d2);
  }
}

}  // namespace

cudaError_t launch_biquad(float *buf, int64_t ss, int64_t cs, int64_t fs, int channels, int n_series, int n_sections,
                          int n_samples, BiquadParams c, float *state, cudaStream_t stream) {
  if (n_series <= 0 || n_samples <= 0)
    return cudaSuccess;
  if (n_sections < 1 || n_sections > kMaxSections)
    return cudaErrorInvalidValue;
  const unsigned grid = (n_series + BQ_SERIES - 1) / BQ_SERIES;
  switch (n_sections) {
    case 1:
      espb_biquad_kernel<1><<<grid, BQ_THREADS, 0, stream>>>(buf, ss, cs, fs, channels, n_series, n_samples, c, state);
      break;
    case 2:
      espb_biquad_kernel<2><<<grid, BQ_THREADS, 0, stream>>>(buf, ss, cs, fs, channels, n_series, n_samples, c, state);
      break;
    case 3:
      espb_biquad_kernel<3><<<grid, BQ_THREADS, 0, stream>>>(buf, ss, cs, fs, channels, n_series, n_samples, c, state);
      break;
    default:
      espb_biquad_kernel<4><<<grid, BQ_THREADS, 0, stream>>>(buf, ss, cs, fs, channels, n_series, n_samples, c, state);
      break;
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace espb
