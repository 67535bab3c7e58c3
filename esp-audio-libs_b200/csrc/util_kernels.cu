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
  const uint64_t stride = (uint64_t) gridDim.x * blockDim.x, tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  // 128-bit loads over the 16-byte aligned middle, words before and after it one by one
  uint64_t head = ((16 - ((uintptr_t) w & 15)) & 15) / 4;
  if (head > n)
    head = n;
  const uint64_t n_vec = (n - head) / 4;
  const uint4 *v = reinterpret_cast<const uint4 *>(w + head);
#pragma unroll 8
  for (uint64_t i = tid; i < n_vec; i += stride) {
    const uint4 q = __ldg(v + i);
    acc += (unsigned long long) q.x + q.y + q.z + q.w;
  }
  for (uint64_t i = tid; i < head; i += stride)
    acc += __ldg(w + i);
  for (uint64_t i = head + n_vec * 4 + tid; i < n; i += stride)
    acc += __ldg(w + i);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    acc += __shfl_xor_sync(0xffffffffu, acc, d);
  // one atomic per CTA (one per warp was 19 k atomics on a single address: 60 % of the kernel's time)
  __shared__ unsigned long long part[8];
  if ((threadIdx.x & 31) == 0)
    part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      t += part[k];
    atomicAdd(sum, t);
  }
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

// The resampler's inner loop on its own: a 4 series x 8 outputs x 2 filters register tile fed from shared memory by
// one per-lane and four warp-uniform 128-bit loads per input row (32 FFMA2 per row), at the kernel's occupancy
// (four 128-thread CTAs of <= 128 registers per SM) — no TMA, no barriers, no epilogue.  What this reaches is the
// practical ceiling of that loop; the distance from it to the FMA-only probes is the cost of the operand pattern
// and of running 4 warps per sub-partition.
constexpr int kTileRows = 32, kTileBpp = 4;
__global__ void __launch_bounds__(128, 4) espb_tile_probe_kernel(float *out, int iters, float seed) {
  extern __shared__ __align__(128) float tile_smem[];
  float *gs = tile_smem;                                // [rows][bpp][16]
  float *xs = tile_smem + kTileRows * kTileBpp * 16;    // [rows][128]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kTileRows * kTileBpp * 16 + kTileRows * 128; i += blockDim.x)
    tile_smem[i] = seed * (float) (i % 7) * 1e-3f;
  __syncthreads();
  float2 acc[4][8];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int n = 0; n < 8; ++n)
      acc[e][n] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int jb = 0; jb < kTileRows / 4; ++jb) {
      const float *xb = xs + lane * 4 + jb * 4 * 128;
      const float *gb = gs + warp * 16 + jb * 4 * kTileBpp * 16;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 xv = *reinterpret_cast<const float4 *>(xb + jj * 128);
        const float4 *gp = reinterpret_cast<const float4 *>(gb + jj * kTileBpp * 16);
        const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
        const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
        const float2 gg[8] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y),
                              make_float2(g1.z, g1.w), make_float2(g2.x, g2.y), make_float2(g2.z, g2.w),
                              make_float2(g3.x, g3.y), make_float2(g3.z, g3.w)};
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            acc[e][n] = __ffma2_rn(gg[n], make_float2(x4[e], x4[e]), acc[e][n]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int n = 0; n < 8; ++n)
      s += acc[e][n].x + acc[e][n].y;
  out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

cudaError_t run_tile_probe(double *tflops) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 4 * 8, iters = 200;
  const size_t smem = (size_t) (kTileRows * kTileBpp * 16 + kTileRows * 128) * sizeof(float) * 2;  // the 2-stage ring
  e = cudaFuncSetAttribute(espb_tile_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(espb_tile_probe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int) cudaSharedmemCarveoutMaxShared);
  float *out = nullptr;
  if (e == cudaSuccess)
    e = cudaMalloc(&out, (size_t) blocks * 128 * sizeof(float));
  if (e != cudaSuccess)
    return e;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  double best = 0.0;
  for (int rep = 0; rep < 5 && e == cudaSuccess; ++rep) {  // first two are warm-up
    cudaEventRecord(t0, 0);
    espb_tile_probe_kernel<<<blocks, 128, smem>>>(out, iters, 1e-3f);
    count_launch();
    cudaEventRecord(t1, 0);
    e = cudaEventSynchronize(t1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flop = 4.0 * 32.0 * 32.0 * kTileRows * (double) iters * blocks * 4.0;  // 32 FFMA2 x 32 lanes x 4 flop
    if (e == cudaSuccess && rep >= 2 && flop / (ms * 1e-3) / 1e12 > best)
      best = flop / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(out);
  if (e == cudaSuccess)
    *tflops = best;
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_checksum(const uint32_t *words, uint64_t n, unsigned long long *sum_dev, cudaStream_t stream) {
  if (n == 0)
    return cudaSuccess;
  uint64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148 * 128)  // (many short grid-stride loops rather than few long ones: pcm_kernels.cu:grid_x_for)
    blocks = 148 * 128;
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
