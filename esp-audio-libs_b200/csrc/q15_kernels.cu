// Q15 mix / volume helpers downstream of the resampler path: the portable-C kernels of the
// reference's dsp.h, batched over device buffers.  Pure HBM-bound int16 work; bit-exact by construction.
//   dsps_add_s16_ansi  (src/dsp/dsps_add_s16_ansi.c:10-27):  out[i*so] = (int16)(((int32) a[i*s1] + b[i*s2]) >> shift)
//   dsps_mulc_s16_ansi (src/dsp/dsps_mulc_s16_ansi.c:18-31): out[i*so] = (int16)(((int32) a[i*si] * C) >> 15)
// Unit-stride, 16-byte aligned buffers take the vector kernels (8 samples per 128-bit load/store, grid-stride);
// anything else the strided ones.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "kernels.hpp"

namespace espb {

namespace {

__device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b, int shift) {
  const int32_t lo = ((int32_t) (int16_t) (a & 0xffffu) + (int32_t) (int16_t) (b & 0xffffu)) >> shift;
  const int32_t hi = ((int32_t) (int16_t) (a >> 16) + (int32_t) (int16_t) (b >> 16)) >> shift;
  return ((uint32_t) lo & 0xffffu) | ((uint32_t) hi << 16);
}

__device__ __forceinline__ uint32_t mulc2(uint32_t a, int32_t c) {
  const int32_t lo = ((int32_t) (int16_t) (a & 0xffffu) * c) >> 15;
  const int32_t hi = ((int32_t) (int16_t) (a >> 16) * c) >> 15;
  return ((uint32_t) lo & 0xffffu) | ((uint32_t) hi << 16);
}

__global__ void __launch_bounds__(256)
    espb_add_s16_vec_kernel(const uint4 *a, const uint4 *b, uint4 *out,
                            uint64_t n_vec, int shift) {
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 x = a[i], y = b[i];  // plain loads: the buffers may alias (in-place mixing, as in the reference)
    out[i] = make_uint4(add2(x.x, y.x, shift), add2(x.y, y.y, shift), add2(x.z, y.z, shift), add2(x.w, y.w, shift));
  }
}

__global__ void __launch_bounds__(256)
    espb_mulc_s16_vec_kernel(const uint4 *a, uint4 *out, uint64_t n_vec, int c) {
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 x = a[i];
    out[i] = make_uint4(mulc2(x.x, c), mulc2(x.y, c), mulc2(x.z, c), mulc2(x.w, c));
  }
}

__global__ void __launch_bounds__(256)
    espb_add_s16_strided_kernel(const int16_t *a, const int16_t *b, int16_t *out,
                                uint64_t first, uint64_t n, int64_t s1, int64_t s2, int64_t so, int shift) {
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  for (uint64_t i = first + (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t acc = (int32_t) a[(int64_t) i * s1] + (int32_t) b[(int64_t) i * s2];
    out[(int64_t) i * so] = (int16_t) (acc >> shift);
  }
}

__global__ void __launch_bounds__(256)
    espb_mulc_s16_strided_kernel(const int16_t *a, int16_t *out, uint64_t first, uint64_t n,
                                 int64_t si, int64_t so, int c) {
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  for (uint64_t i = first + (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[(int64_t) i * so] = (int16_t) (((int32_t) a[(int64_t) i * si] * c) >> 15);
}

int grid_for(uint64_t items) {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // (512 CTAs per SM, i.e. a few grid-stride iterations per thread: 6.9 TB/s against 6.3 with 8 resident CTAs per SM
  //  walking the buffers in one narrow window — see pcm_kernels.cu:grid_x_for)
  static const int per_sm = getenv("ESPB_Q15_CTAS") ? atoi(getenv("ESPB_Q15_CTAS")) : 512;
  const uint64_t want = (items + 255) / 256, cap = (uint64_t) sms * per_sm;
  return (int) (want < cap ? (want ? want : 1) : cap);
}

bool aligned16(const void *p) { return ((uintptr_t) p & 15) == 0; }

}  // namespace

cudaError_t launch_add_s16(const int16_t *a, const int16_t *b, int16_t *out, uint64_t n, int64_t s1, int64_t s2,
                           int64_t so, int shift, cudaStream_t stream) {
  if (n == 0)
    return cudaSuccess;
  uint64_t done = 0;
  if (s1 == 1 && s2 == 1 && so == 1 && aligned16(a) && aligned16(b) && aligned16(out) && n >= 8) {
    const uint64_t n_vec = n / 8;
    espb_add_s16_vec_kernel<<<grid_for(n_vec), 256, 0, stream>>>(reinterpret_cast<const uint4 *>(a),
                                                                  reinterpret_cast<const uint4 *>(b),
                                                                  reinterpret_cast<uint4 *>(out), n_vec, shift);
    count_launch();
    done = n_vec * 8;
  }
  if (done < n) {
    espb_add_s16_strided_kernel<<<grid_for(n - done), 256, 0, stream>>>(a, b, out, done, n, s1, s2, so, shift);
    count_launch();
  }
  return cudaGetLastError();
}

cudaError_t launch_mulc_s16(const int16_t *a, int16_t *out, uint64_t n, int16_t c, int64_t si, int64_t so,
                            cudaStream_t stream) {
  if (n == 0)
    return cudaSuccess;
  uint64_t done = 0;
  if (si == 1 && so == 1 && aligned16(a) && aligned16(out) && n >= 8) {
    const uint64_t n_vec = n / 8;
    espb_mulc_s16_vec_kernel<<<grid_for(n_vec), 256, 0, stream>>>(reinterpret_cast<const uint4 *>(a),
                                                                   reinterpret_cast<uint4 *>(out), n_vec, (int) c);
    count_launch();
    done = n_vec * 8;
  }
  if (done < n) {
    espb_mulc_s16_strided_kernel<<<grid_for(n - done), 256, 0, stream>>>(a, out, done, n, si, so, (int) c);
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace espb
