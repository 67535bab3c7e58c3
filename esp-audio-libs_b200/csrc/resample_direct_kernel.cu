// ART polyphase resampler — direct-input form of the sm_100a kernel (see resample_kernel.cu for the algorithm, the
// expanded coefficient matrix G, the ring / mbarrier protocol and the register tile; this file differs only in where
// the input rows come from).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.hpp"
#include "kernels.hpp"
#include "resample_device.cuh"

namespace espb {

// DIRECT (interleaved stereo float input, no library stage before the resampler): the input rows j >= 0 are not
// staged time-major at all.  A chunk of them is two TMA *tensor* boxes of 64 streams x 16 frames x 2 channels
// (128 bytes per stream) taken straight from the caller's buffer with the 128-byte swizzle, which spreads the 16-byte
// units of the 64 lines over the banks; a lane owns streams (l, l + 32) of the group and reads two stereo frames of
// one stream per LDS.128 — the same five 128-bit loads per row as the time-major form, conflict-free.  Only the
// carried frames (j < 0) still come from the time-major staging buffer, and rows past the input are zero-filled by
// TMA's bounds handling.  The transposing pass over the whole input (7 % of a step) disappears.
template <int BPP, int NST, int CJ, bool EXACT, bool TMCAP, bool DIRECT>
__device__ __forceinline__ void resample_body(const ResampleParams &p, const DirectInput &din, unsigned char *smem_raw) {
  static_assert(!DIRECT || (CJ == 32 && !TMCAP), "direct input: 32-row chunks, caller-layout output");
  constexpr int NTHREADS = BPP * 32;
  constexpr int STAGES = NST;
  constexpr int MAXC = max_chunks_per_cta(BPP, CJ);
  static_assert(kMaxPassesPerCta * BPP * sizeof(int2) <= (size_t) NST * CJ * SGN * sizeof(float), "set-up table");
  constexpr int XS_STAGE = CJ * SGN;                // floats
  constexpr int GS_STAGE = CJ * BPP * kGRowFloats;  // floats
  constexpr uint32_t X_BYTES = XS_STAGE * sizeof(float), G_BYTES = GS_STAGE * sizeof(float);

  float *gs = reinterpret_cast<float *>(smem_raw);                        // [STAGES][CJ][BPP][16]
  float *xs = gs + STAGES * GS_STAGE;                                     // [STAGES][CJ][128]
  uint64_t *full = reinterpret_cast<uint64_t *>(xs + STAGES * XS_STAGE);  // [STAGES] TMA landed
  int *done = reinterpret_cast<int *>(full + STAGES);                     // [2*STAGES] warps done with a stage
  int32_t *jtab = reinterpret_cast<int32_t *>(done + 2 * STAGES);         // [MAXC] first input row of each chunk
  // [MAXC][BPP] what warp w does in chunk c: row groups [r0, r1), end-of-pass flag
  uint16_t *rtab = reinterpret_cast<uint16_t *>(jtab + MAXC);
  // [BPP][NB] schedule entries of the block each warp is finishing, fetched by cp.async during the pass's last chunk
  OutEntry *etab = reinterpret_cast<OutEntry *>(rtab + MAXC * BPP);
  int32_t *hdr = reinterpret_cast<int32_t *>(etab + BPP * NB);  // [4] CTA constants for the refilling lane
  int2 *wtab = reinterpret_cast<int2 *>(smem_raw);  // set-up only (the ring is not in use yet): [MAXP][BPP] windows

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = blockIdx.x;
  const int T = p.taps;

  // ---- which passes / chunks this CTA sweeps
  const int pass_first = p.pass_first + blockIdx.y * p.passes_per_cta;
  int pass_last = pass_first + p.passes_per_cta;
  if (pass_last > p.pass_end)
    pass_last = p.pass_end;
  const int chunk_first = p.pass_chunk_begin[pass_first], chunk_last = p.pass_chunk_begin[pass_last];
  const int n_chunks = __shfl_sync(0xffffffffu, chunk_last - chunk_first, 0);  // warp-uniform by construction

  // ---- build the signal-independent tables this CTA needs (no global loads, no index arithmetic in the main loop)
  for (int i = tid; i < (pass_last - pass_first) * BPP; i += NTHREADS) {  // window [lo, hi) of (pass, warp)
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
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      done[s] = 0;
    }
    hdr[0] = chunk_first - p.g_chunk_base;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  for (int i = tid; i < n_chunks; i += NTHREADS) {
    const ChunkEntry ce = p.chunks[chunk_first + i];
    const bool last_of_pass = (i + 1 == n_chunks) || (p.chunks[chunk_first + i + 1].pass != ce.pass);
    jtab[i] = ce.j_start;
    // rows of this chunk inside each warp's window, in groups of RG (rows outside it only multiply zeros)
#pragma unroll
    for (int w = 0; w < BPP; ++w) {
      const int2 win = wtab[(ce.pass - pass_first) * BPP + w];
      int r0 = win.x - ce.j_start, r1 = win.y - ce.j_start;
      r0 = r0 < 0 ? 0 : (r0 / RG);
      r1 = r1 > CJ ? CJ / RG : ((r1 + RG - 1) / RG);
      if (r0 > CJ / RG)
        r0 = CJ / RG;
      if (r1 < r0)
        r1 = r0;
      if (DIRECT && ce.j_start < 0) {  // carried frames: rows from input frame 0 on belong to the next chunk
        const int lim = (-ce.j_start) / RG;  // (the plan starts these chunks on multiples of RG)
        r1 = r1 < lim ? r1 : lim;
        r0 = r0 < r1 ? r0 : r1;
      }
      rtab[i * BPP + w] = (uint16_t) (r0 | (r1 << 4) | (last_of_pass ? kPassDone : 0) |
                                      ((DIRECT && ce.j_start < 0) ? kHistory : 0));
    }
  }
  __syncthreads();

  // Fill stage c % STAGES with chunk c: two TMA bulk copies (16 KB of G, 16 KB of x) on one mbarrier.
  // (Addresses are rebuilt from the parameters here — one lane runs this once per chunk — rather than held in
  // registers across the FMA loop.)
  auto issue_chunk = [&](int c) {
    const int st = c % STAGES;
    const float *xt_group = p.xt + (int64_t) blockIdx.x * p.xt_rows * SGN;
    mbar_expect_tx(&full[st], X_BYTES + G_BYTES);
    tma_bulk_g2s(gs + st * GS_STAGE, p.G + (size_t) (hdr[0] + c) * GS_STAGE, G_BYTES, &full[st]);
    const int j = jtab[c];
    if (DIRECT && j >= 0) {  // two boxes: frames [j, j+16) and [j+16, j+32) of the group's 64 streams
      tma_tensor2d_g2s(xs + st * XS_STAGE, &din.map, 2 * j, (int) blockIdx.x * (SGN / 2), &full[st]);
      tma_tensor2d_g2s(xs + st * XS_STAGE + XS_STAGE / 2, &din.map, 2 * j + 32, (int) blockIdx.x * (SGN / 2), &full[st]);
    } else {
      tma_bulk_g2s(xs + st * XS_STAGE, xt_group + (int64_t) (j + T) * SGN, X_BYTES, &full[st]);
    }
  };
  if (tid == 0)
    for (int c = 0; c < STAGES && c < n_chunks; ++c)
      issue_chunk(c);

  // accumulators [series e][output n] x (filter 0, filter 1): packed pairs in fast mode, scalars in exact mode
  float2 acc2[EXACT ? 1 : 4][EXACT ? 1 : NB];
  float acc1[EXACT ? 4 : 1][EXACT ? NB : 1][2];
  auto clear_acc = [&]() {
    if constexpr (EXACT) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int n = 0; n < NB; ++n)
          acc1[e][n][0] = acc1[e][n][1] = 0.0f;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int n = 0; n < NB; ++n)
          acc2[e][n] = make_float2(0.0f, 0.0f);
    }
  };
  clear_acc();

  int cur_pass = pass_first;
  uint32_t role_next = rtab[warp];
  for (int c = 0; c < n_chunks; ++c) {
    const int st = c % STAGES;
    const uint32_t role = role_next;
    const bool pass_done = (role & kPassDone) != 0;
    const int r0 = role & 15, r1 = (role >> 4) & 15;
    if (pass_done && lane < NB) {  // the epilogue's schedule entries: global -> shared, no register held meanwhile
      int o = (cur_pass * BPP + warp) * NB + lane;
      o = o < p.n_out ? o : p.n_out - 1;
      cp_async_16(&etab[warp * NB + lane], &p.outs[o]);
    }
    mbar_wait(&full[st], (uint32_t) ((c / STAGES) & 1));
    if constexpr (!DIRECT)
    {
      const float *xrow = xs + st * XS_STAGE + lane * 4;
      const float *grow = gs + st * GS_STAGE + warp * kGRowFloats;
      for (int jb = r0; jb < r1; ++jb) {
        const float *xb = xrow + jb * RG * SGN;
        const float *gb = grow + jb * RG * BPP * kGRowFloats;
#pragma unroll
        for (int jj = 0; jj < RG; ++jj) {
          const float4 xv = *reinterpret_cast<const float4 *>(xb + jj * SGN);
          if constexpr (EXACT) {
            const float4 *gp = reinterpret_cast<const float4 *>(gb + jj * BPP * kGRowFloats);
            const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
            const float g16[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w,
                                   g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc1[e][n][0] = mac<true>(g16[2 * n], x4[e], acc1[e][n][0]);
                acc1[e][n][1] = mac<true>(g16[2 * n + 1], x4[e], acc1[e][n][1]);
              }
          } else {
            const float4 *gp = reinterpret_cast<const float4 *>(gb + jj * BPP * kGRowFloats);
            const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
            const float x4[4] = {xv.x, xv.y, xv.z, xv.w};
            const float2 gg[NB] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y),
                                   make_float2(g1.z, g1.w), make_float2(g2.x, g2.y), make_float2(g2.z, g2.w),
                                   make_float2(g3.x, g3.y), make_float2(g3.z, g3.w)};  // (filter 0, filter 1) pairs
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
              for (int e = 0; e < 4; ++e)
                acc2[e][n] = fma2(gg[n], x4[e], acc2[e][n]);
          }
        }
      }
    }
    else
    {
      // one input row: acc[e][n] (+)= G[row][n][f] * x[e], every accumulator in tap order
      auto row_update = [&](const float (&x4)[4], const float *grow_j) {
        const float4 *gp = reinterpret_cast<const float4 *>(grow_j);
        const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2], g3 = gp[3];
        if constexpr (EXACT) {
          const float g16[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w,
                                 g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
#pragma unroll
          for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc1[e][n][0] = mac<true>(g16[2 * n], x4[e], acc1[e][n][0]);
              acc1[e][n][1] = mac<true>(g16[2 * n + 1], x4[e], acc1[e][n][1]);
            }
        } else {
          const float2 gg[NB] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y),
                                 make_float2(g1.z, g1.w), make_float2(g2.x, g2.y), make_float2(g2.z, g2.w),
                                 make_float2(g3.x, g3.y), make_float2(g3.z, g3.w)};  // (filter 0, filter 1) pairs
#pragma unroll
          for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e)
              acc2[e][n] = fma2(gg[n], x4[e], acc2[e][n]);
        }
      };
      const float *grow = gs + st * GS_STAGE + warp * kGRowFloats;
      if (role & kHistory) {
        // carried frames, time-major in natural series order: streams l and l + 32 are floats [2l, 2l+1] and
        // [64 + 2l, 64 + 2l + 1] of the row (two LDS.64; only the first taps/32 chunks of a call come this way)
        const float *xrow = xs + st * XS_STAGE + lane * 2;
        for (int jb = r0; jb < r1; ++jb) {
          const float *xb = xrow + jb * RG * SGN;
          const float *gb = grow + jb * RG * BPP * kGRowFloats;
#pragma unroll
          for (int jj = 0; jj < RG; ++jj) {
            const float2 xa = *reinterpret_cast<const float2 *>(xb + jj * SGN);
            const float2 xc = *reinterpret_cast<const float2 *>(xb + jj * SGN + SGN / 2);
            const float x4[4] = {xa.x, xa.y, xc.x, xc.y};
            row_update(x4, gb + jj * BPP * kGRowFloats);
          }
        }
      } else {
        // swizzled boxes [half][64 streams][16 frames x 2 ch]: 16-byte unit u of line i sits at unit u ^ (i & 7).
        // 32-bit shared addresses and logic ops only (integer multiply-adds would compete for the FMA pipe).
        const unsigned char *xa = reinterpret_cast<const unsigned char *>(xs + st * XS_STAGE);
        const uint32_t key = ((uint32_t) (lane & 7) << 4) | ((uint32_t) lane << 7);  // line offset | unit key
        for (int jb = r0; jb < r1; ++jb) {
          // rows 4jb .. 4jb+3 = units 2m, 2m + 1 (m = jb & 3) of half jb >> 2
          const uint32_t a0 = ((((uint32_t) jb >> 2) << 13) | (((uint32_t) jb & 3u) << 5)) ^ key;
          const uint32_t a1 = a0 ^ 16u;
          const float *gb = grow + jb * RG * BPP * kGRowFloats;
          // streams l / l + 32: frames (4jb, 4jb+1) and (4jb+2, 4jb+3)
          const float4 va = *reinterpret_cast<const float4 *>(xa + a0);
          const float4 vb = *reinterpret_cast<const float4 *>(xa + a0 + 4096);
          const float4 vc = *reinterpret_cast<const float4 *>(xa + a1);
          const float4 vd = *reinterpret_cast<const float4 *>(xa + a1 + 4096);
          const float x0[4] = {va.x, va.y, vb.x, vb.y}, x1[4] = {va.z, va.w, vb.z, vb.w};
          const float x2[4] = {vc.x, vc.y, vd.x, vd.y}, x3[4] = {vc.z, vc.w, vd.z, vd.w};
          row_update(x0, gb);
          row_update(x1, gb + BPP * kGRowFloats);
          row_update(x2, gb + 2 * BPP * kGRowFloats);
          row_update(x3, gb + 3 * BPP * kGRowFloats);
        }
      }
    }

    // Release the stage.  The last of the BPP warps to get here re-arms it and issues the refill
    // (chunk c + STAGES); nobody waits for anybody.  (A designated refilling warp that waits for the others on an
    // "empty" mbarrier was measured 17 % slower: it cannot run ahead while it waits.)
    role_next = rtab[(c + 1) * BPP + warp];  // (one entry past the CTA's last chunk is still inside the table)
    __syncwarp();
    if (lane == 0) {
      if (smem_arrive(&done[st]) == BPP - 1) {
        done[st] = 0;  // published to the other warps by the release of the mbarrier arrive below
        if (c + STAGES < n_chunks)
          issue_chunk(c + STAGES);
      }
    }

    // ---- end of pass: blend, store, clear
    if (pass_done) {
      cp_async_wait_all();
      __syncwarp();
      const int o0 = (cur_pass * BPP + warp) * NB;
      const OutEntry *et = etab + warp * NB;
      float v[4][NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const OutEntry en = et[n];  // one broadcast LDS.128
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float sum1, sum2;
          if constexpr (EXACT) {
            sum1 = acc1[e][n][0];
            sum2 = acc1[e][n][1];
          } else {
            sum1 = acc2[e][n].x;
            sum2 = acc2[e][n].y;
          }
          if (en.kind == kKindBlend) {  // art_resampler.cpp:450, un-fused
            v[e][n] = __fadd_rn(__fmul_rn(sum2, en.w), __fmul_rn(sum1, __fsub_rn(1.0f, en.w)));
          } else if (en.kind == kKindSingle) {
            v[e][n] = sum1;
          } else {  // pass-through: *source (art_resampler.cpp:426,440) = tap numTaps/2-1 of the window
            const int jsrc = en.ws + T / 2 - 1;
            if (DIRECT && jsrc >= 0) {
              const int64_t stream = (int64_t) group * (SGN / 2) + lane + (e >> 1) * 32;
              v[e][n] = stream * 2 < p.n_series ? din.in[stream * din.in_ss + (int64_t) jsrc * 2 + (e & 1)] : 0.0f;
            } else {
              const int col = DIRECT ? ((e >> 1) * (SGN / 2) + lane * 2 + (e & 1)) : (lane * 4 + e);
              v[e][n] = p.xt[((int64_t) group * p.xt_rows + (jsrc + T)) * SGN + col];
            }
          }
        }
      }
      // series of accumulator row e: natural order 4*lane + e, or (direct input) streams lane and lane + 32
      const int series0 = DIRECT ? group * SGN + lane * 2 : group * SGN + lane * 4;
      auto series_of = [&](int e) { return DIRECT ? series0 + (e >> 1) * (SGN / 2) + (e & 1) : series0 + e; };
      if (TMCAP && p.out_vec == kOutVecTimeMajor) {  // scratch for a following in-library stage: one 16-byte store per lane
        const int row0 = (int) blockIdx.x * (int) p.out_tm_rows + o0;  // < 2^31: the scratch would be > 1 TB otherwise
        float *dst = p.out_tm + (int64_t) row0 * SGN + (threadIdx.x & 31) * 4;
#pragma unroll
        for (int n = 0; n < NB; ++n)
          if (o0 + n < p.n_out)
            *reinterpret_cast<float4 *>(dst + n * SGN) = make_float4(v[0][n], v[1][n], v[2][n], v[3][n]);
      } else if (p.out_vec == kOutVecStereo && o0 + NB <= p.n_out) {
        // interleaved stereo: a lane owns two streams x 8 frames x 2 channels = 2 x 64 contiguous bytes
        float *dst = p.out + (int64_t) (series0 >> 1) * p.out_ss + (int64_t) o0 * 2;
        if (series0 < p.n_series) {
#pragma unroll
          for (int k = 0; k < NB / 2; ++k)
            reinterpret_cast<float4 *>(dst)[k] = make_float4(v[0][2 * k], v[1][2 * k], v[0][2 * k + 1], v[1][2 * k + 1]);
        }
        if (series_of(2) < p.n_series) {
          dst += DIRECT ? 32 * p.out_ss : p.out_ss;
#pragma unroll
          for (int k = 0; k < NB / 2; ++k)
            reinterpret_cast<float4 *>(dst)[k] = make_float4(v[2][2 * k], v[3][2 * k], v[2][2 * k + 1], v[3][2 * k + 1]);
        }
      } else if (!DIRECT && p.out_vec == kOutVecFrame4 && o0 + NB <= p.n_out) {
        // interleaved, channel count a multiple of 4: the lane's 4 series are 16 contiguous bytes of every frame
        if (series0 < p.n_series) {
          const int sidx = series0 / p.channels, ch = series0 - sidx * p.channels;
          float *dst = p.out + (int64_t) sidx * p.out_ss + ch + (int64_t) o0 * p.channels;
#pragma unroll
          for (int n = 0; n < NB; ++n)
            *reinterpret_cast<float4 *>(dst + n * p.channels) = make_float4(v[0][n], v[1][n], v[2][n], v[3][n]);
        }
      } else if (p.out_vec == kOutVecPlanar && o0 + NB <= p.n_out) {
        // frames contiguous per series (planar, or interleaved mono): 8 frames = 32 contiguous bytes per series
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int series = series_of(e);
          if (series < p.n_series) {
            const int sidx = series / p.channels, ch = series - sidx * p.channels;
            float4 *dst = reinterpret_cast<float4 *>(p.out + (int64_t) sidx * p.out_ss + (int64_t) ch * p.out_cs + o0);
            dst[0] = make_float4(v[e][0], v[e][1], v[e][2], v[e][3]);
            dst[1] = make_float4(v[e][4], v[e][5], v[e][6], v[e][7]);
          }
        }
      } else {  // any layout, partial blocks
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int series = series_of(e);
          if (series < p.n_series) {
            const int sidx = series / p.channels, ch = series - sidx * p.channels;
            float *dst = p.out + (int64_t) sidx * p.out_ss + (int64_t) ch * p.out_cs + (int64_t) o0 * p.out_fs;
#pragma unroll
            for (int n = 0; n < NB; ++n)
              if (o0 + n < p.n_out)
                dst[(int64_t) n * p.out_fs] = v[e][n];
          }
        }
      }
      clear_acc();
      ++cur_pass;
    }
  }
}

// The parameters are a __grid_constant__ (the TMA unit reads the tensor map from the parameter space); the ring is
// 1024-byte aligned (the swizzle pattern is address-based).
template <bool EXACT>
__global__ void __launch_bounds__(128, 4) espb_resample_direct_kernel(const __grid_constant__ ResampleDirectParams k) {
  extern __shared__ __align__(1024) unsigned char espb_ring_direct[];
  resample_body<4, 2, 32, EXACT, false, true>(k.p, k.d, espb_ring_direct);
}

template <bool EXACT>
static cudaError_t launch_direct_t(const ResampleParams &p, const DirectInput &d, int n_groups, int n_ctas_y,
                                   cudaStream_t stream) {
  const size_t smem = resample_smem_bytes(4, 32);
  static PerDeviceOnce once;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(espb_resample_direct_kernel<EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int) smem);
    if (e == cudaSuccess)  // four CTAs per SM need the full shared-memory carve-out
      e = cudaFuncSetAttribute(espb_resample_direct_kernel<EXACT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               (int) cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess)
      return e;
    if (getenv("ESPB_DEBUG")) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, espb_resample_direct_kernel<EXACT>, 128, smem);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, espb_resample_direct_kernel<EXACT>);
      fprintf(stderr, "[espb] resample_direct<exact=%d>: smem %zu B, %d regs, occupancy %d CTA/SM\n", (int) EXACT, smem,
              fa.numRegs, nb);
    }
  }
  dim3 grid(n_groups, n_ctas_y);
  ResampleDirectParams k;
  k.p = p;
  k.d = d;
  espb_resample_direct_kernel<EXACT><<<grid, 128, smem, stream>>>(k);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_resample_direct(const ResampleParams &p, const DirectInput &d, int n_groups, int n_ctas_y,
                                   bool exact, cudaStream_t stream) {
  return exact ? launch_direct_t<true>(p, d, n_groups, n_ctas_y, stream)
               : launch_direct_t<false>(p, d, n_groups, n_ctas_y, stream);
}

}  // namespace espb
