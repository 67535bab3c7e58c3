// Stand-alone check of the direct-input staging (not part of the library): one 64-stream x 16-frame stereo box through
// a TMA tensor map with the 128-byte swizzle, map passed inside a __grid_constant__ struct, and the un-swizzling
// address formula the resampler kernel uses.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_box_probe tma_box_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

struct Params {
  float *out;
  int j, n_streams;
  alignas(64) CUtensorMap map;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float *xs = reinterpret_cast<float *>(smem);            // [64 lines][32 floats]
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(8192) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(xs)),
        "l"(&p.map), "r"(2 * p.j), "r"(0), "r"(smem_u32(bar))
        : "memory");
  }
  asm volatile(
      "{\n.reg .pred q;\nW: mbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n@q bra D;\nbra W;\nD:\n}\n" ::"r"(
          smem_u32(bar))
      : "memory");
  // un-swizzle: frame r (0..15), channel c of line i: unit u = r / 2 at unit position u ^ (i & 7)
  for (int k = threadIdx.x; k < 64 * 32; k += blockDim.x) {
    const int i = k / 32, f = k % 32, r = f / 2, c = f % 2;
    const int u = r / 2;
    const unsigned char *base = smem + i * 128 + (((u ^ (i & 7)) << 4));
    p.out[k] = reinterpret_cast<const float *>(base)[(r & 1) * 2 + c];
  }
}

int main() {
  const int n_streams = 5, n_in = 100, stride = 208;  // floats per stream row (multiple of 4)
  std::vector<float> h((size_t) n_streams * stride);
  for (int s = 0; s < n_streams; ++s)
    for (int k = 0; k < stride; ++k)
      h[(size_t) s * stride + k] = s * 1000.0f + k;
  float *d_in, *d_out;
  cudaMalloc(&d_in, h.size() * 4);
  cudaMalloc(&d_out, 64 * 32 * 4);
  cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *ptr = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr);
  Params p{};
  p.out = d_out;
  p.n_streams = n_streams;
  const cuuint64_t dims[2] = {(cuuint64_t) n_in * 2, (cuuint64_t) n_streams};
  const cuuint64_t strides[1] = {(cuuint64_t) stride * 4};
  const cuuint32_t box[2] = {32, 64}, estr[2] = {1, 1};
  CUresult r = ((Fn) ptr)(&p.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_in, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d (entry %p, query %d)\n", (int) r, ptr, (int) qr);
  for (int j : {0, 4, 90}) {
    p.j = j;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 9216);
    probe<<<1, 128, 9216>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(64 * 32);
    cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 64; ++i)
      for (int f = 0; f < 32; ++f) {
        const int col = 2 * j + f;
        const float want = (i < n_streams && col < 2 * n_in) ? i * 1000.0f + col : 0.0f;
        if (o[i * 32 + f] != want && bad++ < 5)
          printf("  j=%d line %d float %d: got %g want %g\n", j, i, f, o[i * 32 + f], want);
      }
    printf("j=%d: %s, mismatches %d\n", j, cudaGetErrorString(e), bad);
  }
  return 0;
}
