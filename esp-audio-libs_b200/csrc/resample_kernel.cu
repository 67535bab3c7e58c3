// ART polyphase resampler — sm_100a kernels.
//
// What is computed (reference art_resampler.cpp:432-451 + dsps_dotprod_f32_ansi.c:17-25):
//   out[series][n] = blend_n( sum_k H[phase_n][k]   * x[series][ws_n + k],
//                             sum_k H[phase_n+1][k] * x[series][ws_n + k] )
// with each sum evaluated strictly in tap order by ONE accumulator (FFMA chain in fast
// mode, FMUL+FADD in exact mode) — splitting a sum across lanes or accumulators moves the
// result by up to 1.8e-6 against the reference and is not allowed.
//
// Mapping.  The schedule (ws_n, phase_n, w_n) is the same for every stream, so the two
// coefficient columns of every output are expanded ONCE per call into a dense,
// pre-skewed matrix G (espb_expand_kernel):
//     G[chunk][row j][block-in-pass b][n in block][f]  =  H[phase_n + f][j - ws_n]  (0 outside the window)
// and the resampler becomes, per input row j, a rank-1 update
//     acc[series e][n][f] += G[j][n][f] * x[j][series e]
// of a 4-series x 8-output x 2-filter register tile (64 independent accumulators per
// thread): x is a per-lane 128-bit shared-memory load (4 series), G sixteen warp-uniform
// values (four broadcast 128-bit loads) -> 64 FFMA per 5 LDS, 8 shared-memory
// wavefronts per 64 FFMA, every accumulator visiting its taps in order j = ws_n .. ws_n+T-1.
//
// A CTA owns 128 series and `BPP` consecutive output blocks (one per warp) — a "pass".
// It sweeps the union of their windows in chunks of 32 input rows through a 3-stage
// shared-memory ring: the G chunk arrives by one bulk-TMA copy (cp.async.bulk + mbarrier
// complete_tx), the x chunk is transposed from the stream-major HBM layout into
// [row][series] by 4-byte cp.async.  Warps whose window does not reach a chunk skip it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.hpp"
#include "kernels.hpp"

namespace espb {

namespace {

constexpr int NB = kOutputsPerBlock;  // 8
constexpr int CJ = kChunkRows;        // 32
constexpr int SGN = kSeriesPerRow;    // 128
constexpr int XROW = SGN + 4;         // padded row: conflict-free transposing stores and 128-bit loads
constexpr int STAGES = 3;
constexpr int MAXC = kMaxChunksPerCta;   // chunk entries cached in shared memory per CTA
constexpr int MAXP = kMaxPassesPerCta;   // passes per CTA

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async_f32(float *dst_smem, const float *src, bool valid) {
  const uint32_t d = smem_u32(dst_smem);
  const int sz = valid ? 4 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool EXACT>
__device__ __forceinline__ float mac(float g, float x, float acc) {
  if (EXACT)
    return __fadd_rn(acc, __fmul_rn(g, x));  // dsps_dotprod_f32_ansi.c:20 — separate multiply and add
  return __fmaf_rn(g, x, acc);
}

}  // namespace

// ---------------------------------------------------------------------------------
// G expansion: one CTA per chunk, one float4 (two outputs x two filters) per thread-iteration.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) espb_expand_kernel(const float *__restrict__ bank,
                                                          const OutEntry *__restrict__ outs,
                                                          const ChunkEntry *__restrict__ chunks, float *__restrict__ G,
                                                          int chunk_first, int n_out, int taps, int bpp) {
  const int gc = chunk_first + blockIdx.x;
  const ChunkEntry ce = chunks[gc];
  const int quads_per_row = bpp * (kGRowFloats / 4);  // float4 per row
  const int total = CJ * quads_per_row;
  float4 *dst = reinterpret_cast<float4 *>(G + (size_t) blockIdx.x * CJ * bpp * kGRowFloats);
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int jj = i / quads_per_row, r = i - jj * quads_per_row;
    const int b = r >> 2, pair = r & 3;  // block in pass, output pair within block
    const int j = ce.j_start + jj;
    float v[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int o = (ce.pass * bpp + b) * NB + pair * 2 + h;
      float c0 = 0.0f, c1 = 0.0f;
      if (o < n_out) {
        const OutEntry e = outs[o];
        const int k = j - e.ws;
        if (k >= 0 && k < taps && e.kind >= kKindSingle) {
          c0 = __ldg(bank + (size_t) e.phase * taps + k);
          if (e.kind == kKindBlend)
            c1 = __ldg(bank + (size_t) (e.phase + 1) * taps + k);
        }
      }
      v[2 * h] = c0;
      v[2 * h + 1] = c1;
    }
    dst[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---------------------------------------------------------------------------------
// Resampler
// ---------------------------------------------------------------------------------
template <int BPP, bool EXACT>
__global__ void __launch_bounds__(BPP * 32, (BPP <= 8 ? 2 : 1))
    espb_resample_kernel(const ResampleParams p) {
  constexpr int QPW = 32 / BPP;  // series quads staged per warp (x4 row groups each)
  static_assert(32 % BPP == 0, "BPP must divide 32");
  constexpr int XS_STAGE = CJ * XROW;             // floats
  constexpr int GS_STAGE = CJ * BPP * kGRowFloats;  // floats
  constexpr uint32_t G_BYTES = GS_STAGE * sizeof(float);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *gs = reinterpret_cast<float *>(smem_raw);               // [STAGES][CJ][BPP][16]
  float *xs = gs + STAGES * GS_STAGE;                            // [STAGES][CJ][XROW]
  uint64_t *gbar = reinterpret_cast<uint64_t *>(xs + STAGES * XS_STAGE);
  ChunkEntry *ctab = reinterpret_cast<ChunkEntry *>(gbar + STAGES);  // [MAXC] this CTA's chunks
  int2 *wtab = reinterpret_cast<int2 *>(ctab + MAXC);                // [MAXP][BPP] window [lo, hi) per (pass, warp)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int series0 = blockIdx.x * SGN;
  const int T = p.taps;

  // ---- which passes / chunks this CTA sweeps
  const int pass_first = p.pass_first + blockIdx.y * p.passes_per_cta;
  int pass_last = pass_first + p.passes_per_cta;
  if (pass_last > p.pass_end)
    pass_last = p.pass_end;
  const int chunk_first = p.pass_chunk_begin[pass_first], chunk_last = p.pass_chunk_begin[pass_last];
  const int n_chunks = chunk_last - chunk_first;

  // ---- cache the signal-independent tables this CTA needs (no dependent global loads in the main loop)
  for (int i = tid; i < n_chunks; i += BPP * 32)
    ctab[i] = p.chunks[chunk_first + i];
  for (int i = tid; i < (pass_last - pass_first) * BPP; i += BPP * 32) {
    const int o0 = ((pass_first + i / BPP) * BPP + (i % BPP)) * NB;
    int2 w = make_int2(0, 0);
    if (o0 < p.n_out) {
      const int o1 = (o0 + NB <= p.n_out ? o0 + NB : p.n_out) - 1;
      w.x = p.outs[o0].ws;
      w.y = p.outs[o1].ws + T;
    }
    wtab[i] = w;
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
      mbar_init(&gbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  // ---- staging role: this thread copies, for QPW series quads, element `se` of the quad
  //      at rows jl, jl+8, jl+16, jl+24 of every chunk
  const int jl = lane >> 2, se = lane & 3;
  const float *src_in[QPW];
  bool series_ok[QPW];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    const int series = series0 + (warp * QPW + q) * 4 + se;
    series_ok[q] = series < p.n_series;
    const int sidx = series_ok[q] ? series : 0;
    const int st = sidx / p.channels, ch = sidx - st * p.channels;
    src_in[q] = p.in + (int64_t) st * p.in_ss + (int64_t) ch * p.in_cs;
  }
  __syncthreads();  // barrier init visible

  auto stage_chunk = [&](int c) {  // c relative to chunk_first
    const int st = c % STAGES;
    const int gc = chunk_first + c;
    const int j0 = ctab[c].j_start;
    if (tid == 0) {
      mbar_expect_tx(&gbar[st], G_BYTES);
      tma_bulk_g2s(gs + st * GS_STAGE, p.G + (size_t) (gc - p.g_chunk_base) * GS_STAGE, G_BYTES, &gbar[st]);
    }
    // this thread's destination: row jl (+8 per step), column of series quad (warp*QPW + q), element se
    float *xdst = xs + st * XS_STAGE + jl * XROW + warp * QPW * 4 + se;
    if ((j0 >= 0) && (j0 + CJ <= p.n_in)) {  // chunk entirely inside this call's input: no per-row tests
      const int64_t joff = (int64_t) (j0 + jl) * p.in_fs;
      const int64_t step = (int64_t) 8 * p.in_fs;
#pragma unroll
      for (int q = 0; q < QPW; ++q) {
        const float *src = src_in[q] + joff;
#pragma unroll
        for (int jb = 0; jb < CJ / 8; ++jb)
          cp_async_f32(xdst + q * 4 + jb * 8 * XROW, src + jb * step, series_ok[q]);
      }
    } else {  // edges: rows before the call come from the carried history, rows past the input are zero
#pragma unroll 1
      for (int q = 0; q < QPW; ++q) {
        const int series = series0 + (warp * QPW + q) * 4 + se;
#pragma unroll 1
        for (int jb = 0; jb < CJ / 8; ++jb) {
          const int j = j0 + jb * 8 + jl;
          const float *src = p.in;
          bool ok = series_ok[q];
          if (j >= 0) {
            ok = ok && (j < p.n_in);
            if (ok)
              src = src_in[q] + (int64_t) j * p.in_fs;
          } else {
            ok = ok && (j >= -T);
            if (ok)
              src = p.hist + (int64_t) series * T + T + j;
          }
          cp_async_f32(xdst + q * 4 + jb * 8 * XROW, src, ok);
        }
      }
    }
  };

  // ---- accumulators: [series e][output n][filter f]
  float acc[4][NB][2];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int n = 0; n < NB; ++n)
      acc[e][n][0] = acc[e][n][1] = 0.0f;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n_chunks)
      stage_chunk(s);
    cp_async_commit();
  }

  int cur_pass = -1, win_lo = 0, win_hi = 0;  // this warp's window [win_lo, win_hi) in input rows
  for (int c = 0; c < n_chunks; ++c) {
    const int st = c % STAGES;
    cp_async_wait<STAGES - 2>();
    mbar_wait(&gbar[st], (uint32_t) ((c / STAGES) & 1));
    __syncthreads();  // every thread's x copies of chunk c have landed; stage (c-1)%STAGES is free
    if (c + STAGES - 1 < n_chunks)
      stage_chunk(c + STAGES - 1);
    cp_async_commit();

    const ChunkEntry ce = ctab[c];
    if (ce.pass != cur_pass) {
      cur_pass = ce.pass;
      const int2 w = wtab[(cur_pass - pass_first) * BPP + warp];
      win_lo = w.x;
      win_hi = w.y;
    }

    // rows of this chunk inside the warp's window, in groups of 8 (rows outside it only multiply zeros)
    int r0 = win_lo - ce.j_start, r1 = win_hi - ce.j_start;
    r0 = r0 < 0 ? 0 : (r0 >> 3);
    r1 = r1 > CJ ? CJ / 8 : ((r1 + 7) >> 3);
    {
      const float *xrow = xs + st * XS_STAGE + lane * 4;
      const float *grow = gs + st * GS_STAGE + warp * kGRowFloats;
      for (int jb = r0; jb < r1; ++jb) {
        const float *xb = xrow + jb * 8 * XROW;
        const float *gb = grow + jb * 8 * BPP * kGRowFloats;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 xv = *reinterpret_cast<const float4 *>(xb + jj * XROW);
          const float4 *gp = reinterpret_cast<const float4 *>(gb + jj * BPP * kGRowFloats);
          const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
          const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
          const float g16[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w,
                                 g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
#pragma unroll
          for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[e][n][0] = mac<EXACT>(g16[2 * n], x4[e], acc[e][n][0]);
              acc[e][n][1] = mac<EXACT>(g16[2 * n + 1], x4[e], acc[e][n][1]);
            }
        }
      }
    }

    // ---- end of pass: blend, store, clear
    const bool pass_done = (c + 1 == n_chunks) || (ctab[c + 1].pass != cur_pass);
    if (pass_done) {
      const int o0 = (cur_pass * BPP + warp) * NB;
      int64_t in_off[4], out_off[4];
      bool live[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int series = series0 + lane * 4 + e;
        live[e] = series < p.n_series;
        const int sidx = series / p.channels, ch = series - sidx * p.channels;
        in_off[e] = (int64_t) sidx * p.in_ss + (int64_t) ch * p.in_cs;
        out_off[e] = (int64_t) sidx * p.out_ss + (int64_t) ch * p.out_cs;
      }
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const int o = o0 + n;
        if (o < p.n_out) {
          const OutEntry en = p.outs[o];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v;
            if (en.kind == kKindBlend) {  // art_resampler.cpp:450, un-fused
              v = __fadd_rn(__fmul_rn(acc[e][n][1], en.w), __fmul_rn(acc[e][n][0], __fsub_rn(1.0f, en.w)));
            } else if (en.kind == kKindSingle) {
              v = acc[e][n][0];
            } else {  // pass-through: *source (art_resampler.cpp:426,440) = tap numTaps/2-1 of the window
              const int j = en.ws + T / 2 - 1;
              v = 0.0f;
              if (live[e]) {
                if (j >= 0)
                  v = (j < p.n_in) ? p.in[in_off[e] + (int64_t) j * p.in_fs] : 0.0f;
                else if (j >= -T)
                  v = p.hist[(int64_t) (series0 + lane * 4 + e) * T + T + j];
              }
            }
            if (live[e])
              p.out[out_off[e] + (int64_t) o * p.out_fs] = v;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int n = 0; n < NB; ++n)
          acc[e][n][0] = acc[e][n][1] = 0.0f;
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// History carry: the last `taps` consumed frames of every series (reference keeps them
// at the front of the ring, art_resampler.cpp:216-222).  hist layout: [series][taps].
// ---------------------------------------------------------------------------------
__global__ void espb_history_kernel(const float *__restrict__ in, int64_t in_ss, int64_t in_cs, int64_t in_fs,
                                    const float *__restrict__ hist_old, float *__restrict__ hist_new, int n_series,
                                    int channels, int taps, int used) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t) n_series * taps)
    return;
  const int series = (int) (i / taps), t = (int) (i - (int64_t) series * taps);
  const int u = used - taps + t;  // frame index relative to this call's input
  float v;
  if (u >= 0) {
    const int st = series / channels, ch = series - st * channels;
    v = in[(int64_t) st * in_ss + (int64_t) ch * in_cs + (int64_t) u * in_fs];
  } else {
    v = (taps + u >= 0) ? hist_old[(int64_t) series * taps + taps + u] : 0.0f;
  }
  hist_new[i] = v;
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
size_t resample_smem_bytes(int bpp) {
  return (size_t) STAGES * (CJ * bpp * kGRowFloats + CJ * XROW) * sizeof(float) + STAGES * sizeof(uint64_t) +
         MAXC * sizeof(ChunkEntry) + (size_t) MAXP * bpp * sizeof(int2);
}

size_t g_chunk_floats(int bpp) { return (size_t) CJ * bpp * kGRowFloats; }

cudaError_t launch_expand(const float *bank, const OutEntry *outs, const ChunkEntry *chunks, float *G,
                          int chunk_first, int n_chunks, int n_out, int taps, int bpp, cudaStream_t stream) {
  if (n_chunks <= 0)
    return cudaSuccess;
  espb_expand_kernel<<<n_chunks, 256, 0, stream>>>(bank, outs, chunks, G, chunk_first, n_out, taps, bpp);
  count_launch();
  return cudaGetLastError();
}

template <int BPP, bool EXACT>
static cudaError_t launch_resample_t(const ResampleParams &p, int n_groups, int n_ctas_y, cudaStream_t stream) {
  const size_t smem = resample_smem_bytes(BPP);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(espb_resample_kernel<BPP, EXACT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess)
      return e;
    configured = true;
  }
  dim3 grid(n_groups, n_ctas_y);
  espb_resample_kernel<BPP, EXACT><<<grid, BPP * 32, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_resample(const ResampleParams &p, int bpp, bool exact, cudaStream_t stream) {
  const int n_groups = (p.n_series + SGN - 1) / SGN;
  const int n_passes = p.pass_end - p.pass_first;
  if (n_groups <= 0 || n_passes <= 0)
    return cudaSuccess;
  const int n_ctas_y = (n_passes + p.passes_per_cta - 1) / p.passes_per_cta;
  if (bpp == 8)
    return exact ? launch_resample_t<8, true>(p, n_groups, n_ctas_y, stream)
                 : launch_resample_t<8, false>(p, n_groups, n_ctas_y, stream);
  if (bpp == 4)
    return exact ? launch_resample_t<4, true>(p, n_groups, n_ctas_y, stream)
                 : launch_resample_t<4, false>(p, n_groups, n_ctas_y, stream);
  return cudaErrorInvalidValue;
}

cudaError_t launch_history(const float *in, int64_t in_ss, int64_t in_cs, int64_t in_fs, const float *hist_old,
                           float *hist_new, int n_series, int channels, int taps, int used, cudaStream_t stream) {
  const int64_t total = (int64_t) n_series * taps;
  if (total <= 0)
    return cudaSuccess;
  const int threads = 256;
  espb_history_kernel<<<(unsigned) ((total + threads - 1) / threads), threads, 0, stream>>>(
      in, in_ss, in_cs, in_fs, hist_old, hist_new, n_series, channels, taps, used);
  count_launch();
  return cudaGetLastError();
}

}  // namespace espb
