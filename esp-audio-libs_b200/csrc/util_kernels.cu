// Small utility kernels: order-independent checksum and the FFMA-only probe that gives
// the roofline denominator for the resampler ("measure an FFMA-only
// microbenchmark on the box and use it as denominator").
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.hpp"

namespace espb {

namespace {

__global__ void __launch_bounds__(256)
    espb_checksum_kernel(const uint32_t *__restrict__ w, uint64_t n, unsigned long long *__restrict__ sum) {
  unsigned long long acc = 0;
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc += __ldg(w + i);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0)
    atomicAdd(sum, acc);
}

// 32 independent FFMA chains per thread, operands in registers, nothing else in the loop.
__global__ void __launch_bounds__(256) espb_fma_probe_kernel(float *out, float a, float b, int iters) {
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k)
    acc[k] = (float) (threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 32; ++k)
      acc[k] = __fmaf_rn(acc[k], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 32; ++k)
    s += acc[k];
  out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same with the packed form (fma.rn.f32x2 -> SASS FFMA2): two FP32 FMAs per lane per instruction.
__global__ void __launch_bounds__(256) espb_fma2_probe_kernel(float *out, float a, float b, int iters) {
  unsigned long long acc[16], aa, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float v = (float) (threadIdx.x + k);
    asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(v));
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[k]) : "l"(aa), "l"(bb));
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[k]));
    s += lo + hi;
  }
  out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

cudaError_t launch_checksum(const uint32_t *words, uint64_t n, unsigned long long *sum_dev, cudaStream_t stream) {
  if (n == 0)
    return cudaSuccess;
  uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 16)
    blocks = 148 * 16;
  espb_checksum_kernel<<<(unsigned) blocks, 256, 0, stream>>>(words, n, sum_dev);
  count_launch();
  return cudaGetLastError();
}

// Best of the scalar (FFMA) and packed (FFMA2) probes; both execute 2*32 flop per thread per iteration.
cudaError_t run_fma_probe(double *tflops, double *clock_mhz, double *tflops_scalar, double *tflops_packed) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256, iters = 16384;
  float *out = nullptr;
  e = cudaMalloc(&out, (size_t) blocks * threads * sizeof(float));
  if (e != cudaSuccess)
    return e;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  double best[2] = {0.0, 0.0};
  for (int which = 0; which < 2 && e == cudaSuccess; ++which) {
    for (int rep = 0; rep < 5; ++rep) {  // first two are warm-up
      cudaEventRecord(t0, 0);
      if (which == 0)
        espb_fma_probe_kernel<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
      else
        espb_fma2_probe_kernel<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
      count_launch();
      cudaEventRecord(t1, 0);
      e = cudaEventSynchronize(t1);
      if (e != cudaSuccess)
        break;
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, t0, t1);
      const double flops = 2.0 * 32.0 * (double) iters * (double) blocks * threads;
      const double tf = flops / (ms * 1e-3) / 1e12;
      if (rep >= 2 && tf > best[which])
        best[which] = tf;
    }
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(out);
  if (e != cudaSuccess)
    return e;
  const double top = best[0] > best[1] ? best[0] : best[1];
  *tflops = top;
  if (tflops_scalar)
    *tflops_scalar = best[0];
  if (tflops_packed)
    *tflops_packed = best[1];
  if (clock_mhz)
    *clock_mhz = top * 1e12 / ((double) sms * 128.0 * 2.0) / 1e6;  // clock that 128 FMA/clk/SM would need
  return cudaGetLastError();
}

}  // namespace espb
