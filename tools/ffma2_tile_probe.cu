// Stand-alone probe (not part of the library): how fast can the resampler's inner-loop pattern run at the kernel's
// occupancy (4 warps of 128 registers per SM sub-partition) when nothing else is in the way?
//   mode 0: 4 series x 8 outputs x 2 filters register tile, operands in registers (32 FFMA2 per "row")
//   mode 1: same, operands fetched from shared memory like the kernel (1 per-lane LDS.128 + 4 broadcast LDS.128 per row)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ffma2_tile_probe ffma2_tile_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int ROWS = 32, BPP = 4;

__device__ __forceinline__ float2 fma2(float2 g, float x, float2 acc) { return __ffma2_rn(g, make_float2(x, x), acc); }

template <int MODE, int RG, int ORDER>
__global__ void __launch_bounds__(128, 4) probe(float *out, int iters, float seed) {
  extern __shared__ __align__(128) float smem[];
  float *gs = smem;                      // [ROWS][BPP][16]
  float *xs = smem + ROWS * BPP * 16;    // [ROWS][128]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ROWS * BPP * 16 + ROWS * 128; i += blockDim.x)
    smem[i] = seed * (float) (i % 7) * 1e-3f;
  __syncthreads();
  float2 acc[4][8];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int n = 0; n < 8; ++n)
      acc[e][n] = make_float2(0.f, 0.f);
  float4 xr = make_float4(seed, seed * 2, seed * 3, seed * 4);
  float4 gr[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    gr[k] = make_float4(seed + k, seed - k, seed * k, seed);
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int jb = 0; jb < ROWS / RG; ++jb) {
      const float *xb = xs + lane * 4 + jb * RG * 128;
      const float *gb = gs + warp * 16 + jb * RG * BPP * 16;
#pragma unroll
      for (int jj = 0; jj < RG; ++jj) {
        float4 xv, g0, g1, g2, g3;
        if (MODE == 1 || MODE == 3) {
          xv = *reinterpret_cast<const float4 *>(xb + jj * 128);
          const float4 *gp = reinterpret_cast<const float4 *>(gb + jj * BPP * 16);
          g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
          if (MODE == 3)  // pin the row order of the FFMA2 blocks (volatile asm statements keep their order)
            asm volatile("" : "+f"(xv.x), "+f"(xv.z), "+f"(g0.x));
        } else {
          xv = xr, g0 = gr[0], g1 = gr[1], g2 = gr[2], g3 = gr[3];
          asm volatile("" : "+f"(xr.x), "+f"(gr[jj & 3].x));  // keep the operands opaque
        }
        const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
        const float2 gg[8] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y),
                              make_float2(g1.z, g1.w), make_float2(g2.x, g2.y), make_float2(g2.z, g2.w),
                              make_float2(g3.x, g3.y), make_float2(g3.z, g3.w)};
        if (MODE == 2) {
#pragma unroll
          for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              acc[e][n] = fma2(gg[0], x4[0], acc[e][n]);
        } else if (ORDER == 0) {
#pragma unroll
          for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              acc[e][n] = fma2(gg[n], x4[e], acc[e][n]);
        } else if (ORDER == 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int n = 0; n < 8; ++n)
              acc[e][n] = fma2(gg[n], x4[e], acc[e][n]);
        } else if (ORDER >= 3 && ORDER <= 7) {
          // pair = two series of one (output, filter): x is the 64-bit operand, the coefficient the broadcast scalar
          const float g16[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w,
                                 g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
          const float2 xp[2] = {make_float2(xv.x, xv.y), make_float2(xv.z, xv.w)};
          float2 *accp = &acc[0][0];  // 32 accumulators viewed as [pair p][k = 2n + f]
          if (ORDER == 3) {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int k = 0; k < 16; ++k)
                accp[p * 16 + k] = __ffma2_rn(xp[p], make_float2(g16[k], g16[k]), accp[p * 16 + k]);
          } else if (ORDER == 5) {  // per G vector (4 scalars): both pairs
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
              for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = kb * 4; k < kb * 4 + 4; ++k)
                  accp[p * 16 + k] = __ffma2_rn(xp[p], make_float2(g16[k], g16[k]), accp[p * 16 + k]);
          } else if (ORDER == 6) {  // per half of the G row (8 scalars): both pairs
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = kb * 8; k < kb * 8 + 8; ++k)
                  accp[p * 16 + k] = __ffma2_rn(xp[p], make_float2(g16[k], g16[k]), accp[p * 16 + k]);
          } else if (ORDER == 7) {  // accumulators laid out [k][p] instead of [p][k]
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int k = 0; k < 16; ++k)
                accp[k * 2 + p] = __ffma2_rn(xp[p], make_float2(g16[k], g16[k]), accp[k * 2 + p]);
          } else if (ORDER == 4) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
#pragma unroll
              for (int p = 0; p < 2; ++p)
                accp[p * 16 + k] = __ffma2_rn(xp[p], make_float2(g16[k], g16[k]), accp[p * 16 + k]);
          }
        } else {  // 2x2 blocks: (n, n+1) x (e, e+1)
#pragma unroll
          for (int n = 0; n < 8; n += 2)
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              acc[e][n] = fma2(gg[n], x4[e], acc[e][n]);
              acc[e + 1][n] = fma2(gg[n], x4[e + 1], acc[e + 1][n]);
              acc[e + 1][n + 1] = fma2(gg[n + 1], x4[e + 1], acc[e + 1][n + 1]);
              acc[e][n + 1] = fma2(gg[n + 1], x4[e], acc[e][n + 1]);
            }
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int n = 0; n < 8; ++n)
      s += acc[e][n].x + acc[e][n].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int RG, int ORDER>
double run(int sms, int ctas_per_sm, int iters) {
  const int blocks = sms * ctas_per_sm * 8;
  const size_t smem = (ROWS * BPP * 16 + ROWS * 128) * sizeof(float) * 2 * (4 / ctas_per_sm);  // same footprint as the kernel's 2-stage ring
  cudaFuncSetAttribute(probe<MODE, RG, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaFuncSetAttribute(probe<MODE, RG, ORDER>, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared);
  float *out;
  cudaMalloc(&out, (size_t) blocks * 128 * sizeof(float));
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(t0);
    probe<MODE, RG, ORDER><<<blocks, 128, smem>>>(out, iters, 1e-3f);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flop = 2.0 * 2.0 * 32.0 * 32.0 * ROWS * (double) iters * blocks * 4.0;  // FFMA2 = 2 FMA = 4 flop per lane
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best)
      best = tf;
  }
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE, RG, ORDER>, 128, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, probe<MODE, RG, ORDER>);
  printf("mode %d rg %d order %d: %.2f TFLOP/s  (%d regs, %d CTA/SM, err=%s)\n", MODE, RG, ORDER, best, fa.numRegs, occ,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  return best;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<1, 4, 0>(sms, 4, 200);
  run<3, 4, 0>(sms, 4, 200);
  run<3, 4, 3>(sms, 4, 200);
  run<3, 8, 3>(sms, 4, 200);
  run<3, 4, 4>(sms, 4, 200);
  run<3, 4, 7>(sms, 4, 200);
  return 0;
}
