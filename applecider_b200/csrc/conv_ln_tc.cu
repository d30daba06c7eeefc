// SpectraNet block front half as ONE tcgen05 kernel (stages whose 3*C_out*groups fit TMEM):
//   three same-padded Conv1d (implicit GEMM, 3-D TMA, per-conv K-block ranges)  ->  + bias
//   ->  LayerNorm over the 3*C concatenated channels of every position  ->  GELU  ->  bf16 [B*L, 3*C]
// The 128 x 384 fp32 accumulator tile lives in TMEM (three 128-column sub-tiles, one per kernel size); an
// accumulator row is one thread of the epilogue, so the LayerNorm statistics are thread-local.
//   stage 1 (C_in 64 -> 3 x 128):  sub-tile j = conv j, one LN group of 384 columns per row.
//   stage 0 (1 -> 3 x 64, polyphase): a GEMM row is 8 positions; one CTA owns two phases, sub-tile j holds
//   [conv j phase r0 | conv j phase r0+1], two LN groups (one per phase) of 3 x 64 columns per row.
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tc;

// optional phase timing (clock64 deltas summed over CTAs): [0] prologue, [1] main loop until the accumulator is ready,
// [2] epilogue, [3] CTA count.  Enabled by acb_debug_timing(1); read + reset by acb_debug_timing_read().
__device__ unsigned long long g_cl_timing[4];
__device__ int g_cl_timing_on = 0;

namespace {

// A pipeline item is (K block, sub-tile): A tile + ONE 128-row weight sub-tile (32 KB), 5 stages deep.  K blocks
// where several kernel sizes are active (12 % of them) re-load the A tile once per active sub-tile.
constexpr int CL_STAGES = 5;
constexpr uint32_t CL_STG_BYTES = 32 * 144;  // per-warp 32 x 64 bf16 transpose tile (row pitch 128 + 16 bytes)
constexpr uint32_t CL_A_BYTES = TC_BM * TC_BK * 2;       // 16 KB
constexpr uint32_t CL_SUB_BYTES = 128 * TC_BK * 2;       // 16 KB per 128-row weight sub-tile
constexpr uint32_t CL_STAGE_BYTES = CL_A_BYTES + CL_SUB_BYTES;

struct ConvLnArgs {
  int Lbox, Bbox, tps, nbatch, L;     // A tile geometry (as in gemm_tc)
  int taps, pad, cpt, Cin;
  int kb_lo[3], kb_hi[3];             // K-block range of every sub-tile (nested in sub-tile 2's range)
  int brow_base[3], brow_stride_y;    // packed-weight row of sub-tile j for blockIdx.y
  int ng;                             // LayerNorm groups per accumulator row (1 | 2); group width gw = 128 / ng per sub-tile
  long long row_mul, row_add_y;       // output row = m * row_mul + blockIdx.y * row_add_y + g
  long long out_rows;                 // valid output rows (positions)
  int ldc;                            // 3 * gw
  const float* bias;                  // packed like the weight rows
  const float* gamma;                 // [3*gw] LayerNorm weight (channel order conv0|conv1|conv2)
  const float* beta;
  float eps;
  bf16* out;
  // HANKEL mode (stage 0, polyphase): the A operand is never materialised.  A(m', chunk c) = W16[m' + c] in 16-byte
  // units of the zero-padded signal, so the UMMA descriptor (K-major, no swizzle: rows 16 B apart, SBO 128 B, LBO 16 B)
  // reads overlapping core matrices straight from one 4 KB window of the signal in shared memory.
  const bf16* xwin;            // padded signal [nbatch, a_batch_stride]
  long long xwin_stride;       // elements per sample
  long long xwin_total;        // elements in the whole buffer
  // fused 1x1 downsample + pair max (persistent stage-0 kernel only): pairmax[(4*m + y), 0..63] = max over the CTA's two
  // phases of (LN/GELU row) x Wd^T + bd
  const float* down_bias;      // [64]
  bf16* down_out;              // [out_rows / 2, 64]
};

// ---- epilogue of one tile: bias -> LayerNorm(3*gw) -> GELU -> bf16, executed by the 8 epilogue warps ---------------
// Sub-tiles 0 and 1 sit at TMEM columns 0 and 128; sub-tile 2 at col2 (256, or 384 in the persistent kernel's odd tiles).
// DOWN (persistent stage-0 kernel): the normalised bf16 rows are not stored; they are written into a 128 x 192 K-major
// SWIZZLE_128B tile (a2) and multiplied with the 1x1 downsample weights (wd_smem) by 12 more UMMAs per phase into the TMEM
// columns the phase's conv-0 accumulators have just vacated; the two phases are max-ed and stored as one "pair max" row.
struct ClDown {
  uint32_t a2, wd;        // shared-memory addresses (1024-byte aligned)
  uint32_t bar_d2;        // MMA-complete barrier
  uint32_t* d2_phase;     // per-thread phase counter of bar_d2
};

template <int NPQ, bool DOWN = false>  // NPQ epilogue warps per TMEM lane quarter: 64-column blocks are dealt round-robin to them
__device__ __forceinline__ void cl_epilogue_tile(const ConvLnArgs& p, int mt, int y, uint32_t tmem_base, uint32_t col2, const float* s_bias,
                                                 const float* s_gamma, const float* s_beta, float (*part)[NPQ][128][2], uint8_t* stg,
                                                 int warp, int lane, const ClDown* dn = nullptr) {
  const int sample0 = (mt / p.tps) * p.Bbox;
  const int l0 = (mt % p.tps) * p.Lbox;
  {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // which of the NPQ warps serving this lane quarter
    const int r = q * 32 + lane;
    const int s_in_tile = r / p.Lbox;
    const int l = l0 + (r - s_in_tile * p.Lbox);
    const int sample = sample0 + s_in_tile;
    const long long m = (long long)sample * p.L + l;
    const bool valid_row = (s_in_tile < p.Bbox) && (sample < p.nbatch) && (l < p.L);
    const int gw = 128 / p.ng;
    const float inv_n = 1.0f / (float)(3 * gw);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float d2keep[32];  // DOWN: phase-0 result of this thread's row (32 of the 64 output channels)
    for (int g = 0; g < p.ng; ++g) {
      const long long orow = m * p.row_mul + (long long)y * p.row_add_y + g;
      const bool valid = valid_row && orow < p.out_rows;
      // pass 1: statistics of (acc + bias) over the 3*gw channels of this position
      float2 sum2 = make_float2(0.0f, 0.0f), sq2 = make_float2(0.0f, 0.0f);  // packed fp32: even / odd channels
      for (int j = 0; j < 3; ++j) {
        for (int c0 = 0; c0 < gw; c0 += 32) {
          if ((((g * 3 * gw + j * gw + c0) >> 6) % NPQ) != half) continue;  // 64-column blocks are dealt round-robin to the warps
          uint32_t raw[32];
          tmem_ld32(lane_base + (j == 2 ? col2 : (uint32_t)(128 * j)) + (uint32_t)(g * gw + c0), raw);
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + 128 * j + g * gw + c0);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 bb = b4[i4];
            const float2 v01 = __fadd2_rn(make_float2(__uint_as_float(raw[i4 * 4 + 0]), __uint_as_float(raw[i4 * 4 + 1])), make_float2(bb.x, bb.y));
            const float2 v23 = __fadd2_rn(make_float2(__uint_as_float(raw[i4 * 4 + 2]), __uint_as_float(raw[i4 * 4 + 3])), make_float2(bb.z, bb.w));
            sum2 = __fadd2_rn(sum2, __fadd2_rn(v01, v23));
            sq2 = __ffma2_rn(v23, v23, __ffma2_rn(v01, v01, sq2));
          }
        }
      }
      float sum = sum2.x + sum2.y, sq = sq2.x + sq2.y;
      part[g][half][r][0] = sum;
      part[g][half][r][1] = sq;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * NPQ) : "memory");  // the warps of this quarter
#pragma unroll
      for (int o = 1; o < NPQ; ++o) {
        const int oh = (half + o) % NPQ;
        sum += part[g][oh][r][0];
        sq += part[g][oh][r][1];
      }
      const float mean = sum * inv_n;
      const float rstd = rsqrtf(fmaxf(sq * inv_n - mean * mean, 0.0f) + p.eps);
      // pass 2: normalise, affine, GELU, pack to bf16; 32 x 64 tiles go through a per-warp smem transpose so that every
      // store instruction writes four full 128-byte lines (uncoalesced per-row stores cost 32 L1 wavefronts each)
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      for (int j = 0; j < 3; ++j) {
        for (int cb = 0; cb < gw; cb += 64) {
          if ((((g * 3 * gw + j * gw + cb) >> 6) % NPQ) != half) continue;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int c0 = cb + hh * 32;
            uint32_t raw[32];
            tmem_ld32(lane_base + (j == 2 ? col2 : (uint32_t)(128 * j)) + (uint32_t)(g * gw + c0), raw);
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + 128 * j + g * gw + c0);
            const float4* g4 = reinterpret_cast<const float4*>(s_gamma + 128 * j + g * gw + c0);
            const float4* e4 = reinterpret_cast<const float4*>(s_beta + 128 * j + g * gw + c0);
            uint32_t pk[16];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 bb = b4[i4], gg = g4[i4], ee = e4[i4];
              // packed fp32 (two channels per instruction): ((raw + bias) - mean) * rstd * gamma + beta, then GELU
              const float2 nm2 = make_float2(-mean, -mean), rs2 = make_float2(rstd, rstd);
              const float2 a01 = __fmul2_rn(__fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(raw[i4 * 4 + 0]), __uint_as_float(raw[i4 * 4 + 1])),
                                                                  make_float2(bb.x, bb.y)), nm2), rs2);
              const float2 a23 = __fmul2_rn(__fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(raw[i4 * 4 + 2]), __uint_as_float(raw[i4 * 4 + 3])),
                                                                  make_float2(bb.z, bb.w)), nm2), rs2);
              const float2 y01 = gelu_bf16x2(__ffma2_rn(a01, make_float2(gg.x, gg.y), make_float2(ee.x, ee.y)));
              const float2 y23 = gelu_bf16x2(__ffma2_rn(a23, make_float2(gg.z, gg.w), make_float2(ee.z, ee.w)));
              __nv_bfloat162 h0 = __floats2bfloat162_rn(y01.x, y01.y), h1 = __floats2bfloat162_rn(y23.x, y23.y);
              pk[i4 * 2 + 0] = *reinterpret_cast<uint32_t*>(&h0);
              pk[i4 * 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
            }
            if constexpr (DOWN) {
              // K-major SWIZZLE_128B A tile: K block j (64 channels = 128 bytes per row), 16-byte chunk c of row r at c ^ (r & 7)
#pragma unroll
              for (int v4 = 0; v4 < 4; ++v4) {
                const uint32_t addr = dn->a2 + (uint32_t)j * 16384u + (uint32_t)r * 128u + ((uint32_t)((hh * 4 + v4) ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[v4 * 4 + 0]), "r"(pk[v4 * 4 + 1]), "r"(pk[v4 * 4 + 2]),
                             "r"(pk[v4 * 4 + 3])
                             : "memory");
              }
              continue;
            }
            uint4* srow = reinterpret_cast<uint4*>(stg + lane * 144 + hh * 64);
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) srow[v4] = make_uint4(pk[v4 * 4 + 0], pk[v4 * 4 + 1], pk[v4 * 4 + 2], pk[v4 * 4 + 3]);
          }
          if constexpr (DOWN) continue;
          __syncwarp();
          const int ch = j * gw + cb;  // first channel of this 64-wide block inside the concatenated row
#pragma unroll
          for (int it = 0; it < 8; ++it) {  // 8 lanes x 16 B = one 128-byte row segment, 4 rows per instruction
            const int rr = it * 4 + (lane >> 3), cg = lane & 7;
            const long long orr = __shfl_sync(0xffffffffu, orow, rr);
            if ((vmask >> rr) & 1u) {
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 144 + cg * 16);
              *reinterpret_cast<uint4*>(p.out + orr * p.ldc + ch + cg * 8) = val;
            }
          }
          __syncwarp();
        }
      }
      if constexpr (DOWN) {
        // all 8 warps have written their share of the 128 x 192 tile (and are done with this phase's conv-0 columns)
        fence_proxy_async();
        tc_fence_before();
        asm volatile("bar.sync 10, 256;" ::: "memory");
        const uint32_t d2col = (uint32_t)(g * 64);  // TMEM columns [64 g, 64 g + 64): phase g of sub-tile 0, now dead
        if (warp == 2 && elect_one_sync()) {
          constexpr uint32_t IDESC_D = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) {
            const uint64_t da = make_smem_desc(dn->a2 + kb * 16384u), db = make_smem_desc(dn->wd + kb * 8192u);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_base + d2col, da + 2 * k, db + 2 * k, IDESC_D, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(dn->bar_d2);
        }
        __syncwarp();
        mbar_wait(dn->bar_d2, *dn->d2_phase & 1u);
        *dn->d2_phase += 1u;
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld32(lane_base + d2col + (uint32_t)(half * 32), raw);
        if (g == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) d2keep[i] = __uint_as_float(raw[i]);
        } else {
          const long long prow = 4 * m + y;  // pair index: positions 8 m + 2 y and 8 m + 2 y + 1
          if (valid_row && 2 * prow < p.out_rows) {
            const float4* b4 = reinterpret_cast<const float4*>(p.down_bias + half * 32);
            uint4* dst = reinterpret_cast<uint4*>(p.down_out + prow * 64 + half * 32);
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float4 bb = __ldg(b4 + v4 * 2 + e);
                const int i0 = v4 * 8 + e * 4;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(d2keep[i0], __uint_as_float(raw[i0])) + bb.x,
                                                          fmaxf(d2keep[i0 + 1], __uint_as_float(raw[i0 + 1])) + bb.y);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(d2keep[i0 + 2], __uint_as_float(raw[i0 + 2])) + bb.z,
                                                          fmaxf(d2keep[i0 + 3], __uint_as_float(raw[i0 + 3])) + bb.w);
                w4[e * 2] = *reinterpret_cast<uint32_t*>(&h0);
                w4[e * 2 + 1] = *reinterpret_cast<uint32_t*>(&h1);
              }
              dst[v4] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
          }
        }
      }
    }
  }
}

constexpr int CL_THREADS = 64 + 256;  // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)

template <bool HANKEL>
__global__ void __launch_bounds__(CL_THREADS, 1) conv_ln_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ ConvLnArgs p) {
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  constexpr int NST = HANKEL ? 2 * CL_STAGES : CL_STAGES;            // HANKEL stages hold only a 16 KB weight sub-tile
  constexpr uint32_t STAGE_BYTES = HANKEL ? CL_SUB_BYTES : CL_STAGE_BYTES;
  constexpr uint32_t WIN_BYTES = (128 + 8 * 17 + 8) * 16;            // 128 rows + 17 K blocks of 8 chunks (+ slack)
  __shared__ __align__(8) uint64_t bars[2 * NST + 2];
  __shared__ uint32_t tmem_holder;
  __shared__ float part[2][2][128][2];  // [group][half][row][sum, sumsq] partial LayerNorm statistics
  __shared__ __align__(16) float s_bias[384], s_gamma[384], s_beta[384];  // staged once per CTA while the main loop runs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_start = clock64();
  const int mt = blockIdx.x;
  const int sample0 = (mt / p.tps) * p.Bbox;
  const int l0 = (mt % p.tps) * p.Lbox;
  const int kb_lo = p.kb_lo[2], kb_hi = p.kb_hi[2];
  const int nkb = kb_hi - kb_lo;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[NST]);
  const uint32_t bar_acc = smem_u32(&bars[2 * NST]);
  const uint32_t bar_win = smem_u32(&bars[2 * NST + 1]);
  // smem carve-up: [pipeline stages][8 per-warp transpose tiles][signal window (HANKEL)]
  const uint32_t stg_off = (uint32_t)CL_STAGES * CL_STAGE_BYTES;  // both modes use the same pipeline footprint
  const uint32_t win_off = stg_off + 8 * CL_STG_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_win, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;
  const long long t_pro = clock64();

  if (warp == 0) {
    const uint32_t a_box_bytes = (uint32_t)TC_BK * 2u * (uint32_t)p.Lbox * (uint32_t)p.Bbox;
    if (HANKEL && elect_one_sync()) {
      const long long e0 = (long long)sample0 * p.xwin_stride + 8LL * l0;  // first element of the window
      long long nbytes = (p.xwin_total - e0) * 2;
      if (nbytes > (long long)WIN_BYTES) nbytes = WIN_BYTES;
      nbytes &= ~15LL;
      mbar_expect_tx(bar_win, (uint32_t)nbytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_base + win_off),
                   "l"(p.xwin + e0), "r"((uint32_t)nbytes), "r"(bar_win)
                   : "memory");
    }
    __syncwarp();
    int it = 0;
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      const int tap = kb / p.cpt, cc = kb - tap * p.cpt;
      for (int j = 2; j >= 0; --j) {
        if (kb < p.kb_lo[j] || kb >= p.kb_hi[j]) continue;
        const int s = it % NST;
        const uint32_t ph = (uint32_t)(it / NST) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          if (HANKEL) {
            mbar_expect_tx(bar_full + 8 * s, CL_SUB_BYTES);
            tma_load_2d(sa, &tmB, tap * p.Cin + cc * TC_BK, p.brow_base[j] + (int)blockIdx.y * p.brow_stride_y, bar_full + 8 * s);
          } else {
            mbar_expect_tx(bar_full + 8 * s, a_box_bytes + CL_SUB_BYTES);
            tma_load_3d(sa, &tmA, cc * TC_BK, l0 + tap - p.pad, sample0, bar_full + 8 * s);
            tma_load_2d(sa + CL_A_BYTES, &tmB, tap * p.Cin + cc * TC_BK, p.brow_base[j] + (int)blockIdx.y * p.brow_stride_y, bar_full + 8 * s);
          }
        }
        __syncwarp();
        ++it;
      }
    }
  } else if (warp == 1) {
    if (HANKEL) mbar_wait(bar_win, 0);
    int it = 0;
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      for (int j = 2; j >= 0; --j) {
        if (kb < p.kb_lo[j] || kb >= p.kb_hi[j]) continue;
        const int s = it % NST;
        const uint32_t ph = (uint32_t)(it / NST) & 1u;
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint32_t sb = HANKEL ? sa : sa + CL_A_BYTES;
          // HANKEL: chunk index of MMA k = 8*kb + 2*k; no-swizzle K-major descriptor, LBO = 16 B, SBO = 128 B
          const uint64_t da = HANKEL ? ((uint64_t)(((smem_base + win_off + (uint32_t)(8 * kb) * 16u) & 0x3FFFFu) >> 4) | (1ull << 16) | (8ull << 32) | (1ull << 46))
                                     : make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(tmem_base + 128u * j, da + 2 * k, db + 2 * k, IDESC, (kb > p.kb_lo[j] || k > 0) ? 1u : 0u);
          umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
        ++it;
      }
    }
    if (elect_one_sync()) umma_commit(bar_acc);
    __syncwarp();
  } else {
    // ===================== epilogue: bias -> LayerNorm(3*gw) -> GELU -> bf16 =====================
    {  // stage the per-channel parameters (the epilogue warps are idle during the main loop)
      const int gw0 = 128 / p.ng;
      for (int i = threadIdx.x - 64; i < 384; i += 256) {
        const int j = i >> 7, c = i & 127;  // accumulator column 128*j + c  <->  (group g = c / gw, channel c % gw)
        s_bias[i] = __ldg(p.bias + p.brow_base[j] + (int)blockIdx.y * p.brow_stride_y + c);
        const int g = c / gw0, cc = c - g * gw0;
        s_gamma[i] = __ldg(p.gamma + j * gw0 + cc);
        s_beta[i] = __ldg(p.beta + j * gw0 + cc);
      }
      asm volatile("bar.sync 9, 256;" ::: "memory");  // the 8 epilogue warps
    }
    mbar_wait_sleep(bar_acc, 0);
    tc_fence_after();
    const long long t_acc = clock64();
    cl_epilogue_tile<2>(p, mt, (int)blockIdx.y, tmem_base, 256u, s_bias, s_gamma, s_beta, part,
                     smem_raw + (smem_base - smem_u32(smem_raw)) + stg_off + (size_t)(warp - 2) * CL_STG_BYTES, warp, lane);
    if (g_cl_timing_on && warp == 2 && lane == 0) {
      const long long t_end = clock64();
      atomicAdd(&g_cl_timing[0], (unsigned long long)(t_pro - t_start));
      atomicAdd(&g_cl_timing[1], (unsigned long long)(t_acc - t_pro));
      atomicAdd(&g_cl_timing[2], (unsigned long long)(t_end - t_acc));
      atomicAdd(&g_cl_timing[3], 1ull);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- persistent stage-0 (HANKEL) kernel with a ROTATING TMEM layout ------------------------------------------------
// One CTA per SM walks the tiles (signal window mt, phase pair y).  A tile needs 384 accumulator columns, so TMEM
// (512 columns) cannot hold two tiles -- but the k=1021 conv (sub-tile 2) is 17 of the 20 K blocks of a tile, and the
// 128 spare columns can hold ITS next accumulator: sub-tile 2 alternates between columns [256,384) and [384,512).
// The MMA warp therefore runs sub-tile 2 of tile n+1 while the 8 epilogue warps still normalise tile n, waits for
// them, and only then issues the 3 short K blocks of sub-tiles 1 and 0 into the shared columns [0,256).
// Steady state per tile ~ max(main loop, epilogue) instead of their sum, and no per-tile CTA set-up.
__device__ __forceinline__ void cl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int CLP_NST = 2 * CL_STAGES;                       // 16 KB weight sub-tile stages (8 with the fused downsample)
constexpr int CLP_NPQ = 2;                                   // epilogue warps per TMEM lane quarter (3 was measured: no gain, and
                                                             // 2 keeps the statistics' summation order of the per-tile kernel)
constexpr int CLP_THREADS = 64 + 128 * CLP_NPQ;
constexpr uint32_t CLP_WIN_BYTES = (128 + 8 * 17 + 8) * 16;  // 4352 B signal window

template <bool DOWN>
__global__ void __launch_bounds__(CLP_THREADS, 1) conv_ln_hankel_persist_kernel(const __grid_constant__ CUtensorMap tmB,
                                                                               const __grid_constant__ CUtensorMap tmD,
                                                                               const __grid_constant__ ConvLnArgs p, int MT, int NY) {
  constexpr int NST = DOWN ? 8 : CLP_NST;
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * CLP_NST + 8];
  __shared__ uint32_t tmem_holder;
  __shared__ float part[2][CLP_NPQ][128][2];
  __shared__ __align__(16) float s_bias[4][384], s_gamma[384], s_beta[384];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[CLP_NST]);
  const uint32_t bar_wd = smem_u32(&bars[2 * CLP_NST + 6]);       // downsample weights landed
  const uint32_t bar_d2 = smem_u32(&bars[2 * CLP_NST + 7]);       // downsample MMAs of one phase complete
  const uint32_t bar_acc = smem_u32(&bars[2 * CLP_NST]);           // accumulators of a tile complete (MMA -> epilogue)
  const uint32_t bar_epi = smem_u32(&bars[2 * CLP_NST + 1]);       // epilogue has drained a tile (8 warps -> MMA)
  const uint32_t bar_win_full = smem_u32(&bars[2 * CLP_NST + 2]);  // [2] signal window landed
  const uint32_t bar_win_free = smem_u32(&bars[2 * CLP_NST + 4]);  // [2] all MMAs reading the window have retired
  // smem: [weight stages][transpose tiles | DOWN: 48 KB A2 tile + 24 KB downsample weights][2 signal windows]
  const uint32_t stg_off = (uint32_t)NST * CL_SUB_BYTES;
  const uint32_t win_off = stg_off + (DOWN ? 49152u + 24576u : 4 * CLP_NPQ * CL_STG_BYTES);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_wd, 1);
    mbar_init(bar_d2, 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_epi, 4 * CLP_NPQ);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_win_full + 8 * b, 1);
      mbar_init(bar_win_free + 8 * b, 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    // ===================== producer: signal windows + weight sub-tiles in MMA order =====================
    if (DOWN && elect_one_sync()) {  // the 64 x 192 downsample weights, once per CTA: three 64 x 64 K-major SWIZZLE_128B blocks
      mbar_expect_tx(bar_wd, 24576u);
      for (int kb = 0; kb < 3; ++kb) tma_load_2d(smem_base + stg_off + 49152u + kb * 8192u, &tmD, kb * TC_BK, 0, bar_wd);
    }
    __syncwarp();
    int it = 0, wi = 0;
    for (int mt = blockIdx.x; mt < MT; mt += gridDim.x, ++wi) {
      const int wb = wi & 1;
      mbar_wait(bar_win_free + 8 * wb, (((uint32_t)wi >> 1) & 1u) ^ 1u);
      if (elect_one_sync()) {
        const int sample0 = (mt / p.tps) * p.Bbox, l0 = (mt % p.tps) * p.Lbox;
        const long long e0 = (long long)sample0 * p.xwin_stride + 8LL * l0;
        long long nbytes = (p.xwin_total - e0) * 2;
        if (nbytes > (long long)CLP_WIN_BYTES) nbytes = CLP_WIN_BYTES;
        nbytes &= ~15LL;
        mbar_expect_tx(bar_win_full + 8 * wb, (uint32_t)nbytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_base + win_off + wb * CLP_WIN_BYTES),
                     "l"(p.xwin + e0), "r"((uint32_t)nbytes), "r"(bar_win_full + 8 * wb)
                     : "memory");
      }
      __syncwarp();
      for (int y = 0; y < NY; ++y) {
        for (int jj = 0; jj < 3; ++jj) {
          const int j = 2 - jj;  // sub-tile 2 first (it overlaps the previous tile's epilogue)
          for (int kb = p.kb_lo[j]; kb < p.kb_hi[j]; ++kb, ++it) {
            const int s = it % NST;
            const uint32_t ph = (uint32_t)(it / NST) & 1u;
            mbar_wait_sleep(bar_empty + 8 * s, ph ^ 1u);  // sleeps: a spinning warp steals issue slots from the epilogue warps
            if (elect_one_sync()) {
              mbar_expect_tx(bar_full + 8 * s, CL_SUB_BYTES);
              tma_load_2d(smem_base + s * CL_SUB_BYTES, &tmB, kb * TC_BK, p.brow_base[j] + y * p.brow_stride_y, bar_full + 8 * s);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int it = 0, wi = 0, n = 0;
    for (int mt = blockIdx.x; mt < MT; mt += gridDim.x, ++wi) {
      const int wb = wi & 1;
      mbar_wait(bar_win_full + 8 * wb, ((uint32_t)wi >> 1) & 1u);
      const uint32_t win = smem_base + win_off + wb * CLP_WIN_BYTES;
      for (int y = 0; y < NY; ++y, ++n) {
        for (int jj = 0; jj < 3; ++jj) {
          const int j = 2 - jj;
          if (jj == 1 && n > 0) {  // columns [0,256) are shared with the previous tile: wait for its epilogue
            mbar_wait_sleep(bar_epi, ((uint32_t)(n - 1)) & 1u);
            tc_fence_after();
          }
          const uint32_t acc = tmem_base + (j == 2 ? 256u + 128u * (uint32_t)(n & 1) : 128u * (uint32_t)j);
          for (int kb = p.kb_lo[j]; kb < p.kb_hi[j]; ++kb, ++it) {
            const int s = it % NST;
            const uint32_t ph = (uint32_t)(it / NST) & 1u;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint64_t da = (uint64_t)(((win + (uint32_t)(8 * kb) * 16u) & 0x3FFFFu) >> 4) | (1ull << 16) | (8ull << 32) | (1ull << 46);
              const uint64_t db = make_smem_desc(smem_base + s * CL_SUB_BYTES);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, IDESC, (kb > p.kb_lo[j] || k > 0) ? 1u : 0u);
              umma_commit(bar_empty + 8 * s);
            }
            __syncwarp();
          }
        }
        if (elect_one_sync()) {
          umma_commit(bar_acc);
          if (y == NY - 1) umma_commit(bar_win_free + 8 * wb);  // the window may be overwritten once these MMAs retire
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps =====================
    {
      const int gw0 = 128 / p.ng;
      for (int i = threadIdx.x - 64; i < 384; i += 128 * CLP_NPQ) {
        const int j = i >> 7, c = i & 127;
        for (int y = 0; y < NY; ++y) s_bias[y][i] = __ldg(p.bias + p.brow_base[j] + y * p.brow_stride_y + c);
        const int g = c / gw0, cc = c - g * gw0;
        s_gamma[i] = __ldg(p.gamma + j * gw0 + cc);
        s_beta[i] = __ldg(p.beta + j * gw0 + cc);
      }
      asm volatile("bar.sync 9, %0;" ::"r"(128 * CLP_NPQ) : "memory");
    }
    uint8_t* stg = smem_raw + (smem_base - smem_u32(smem_raw)) + stg_off + (size_t)(warp - 2) * CL_STG_BYTES;
    uint32_t d2_phase = 0;
    ClDown dn;
    dn.a2 = smem_base + stg_off;
    dn.wd = smem_base + stg_off + 49152u;
    dn.bar_d2 = bar_d2;
    dn.d2_phase = &d2_phase;
    if (DOWN) mbar_wait_sleep(bar_wd, 0);
    int n = 0;
    for (int mt = blockIdx.x; mt < MT; mt += gridDim.x) {
      for (int y = 0; y < NY; ++y, ++n) {
        mbar_wait_sleep(bar_acc, (uint32_t)n & 1u);
        tc_fence_after();
        cl_epilogue_tile<CLP_NPQ, DOWN>(p, mt, y, tmem_base, 256u + 128u * (uint32_t)(n & 1), s_bias[y], s_gamma, s_beta, part, stg, warp, lane, &dn);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) cl_mbar_arrive(bar_epi);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- persistent kernel for the tap form (stage 1: C_in = 64 -> 3 x 128, k = 3/31/251), same rotating TMEM layout ------
// Sub-tile 2 (k = 251) is 251 of the 285 K blocks of a tile, so the whole LayerNorm + GELU epilogue of tile n hides
// behind it; the separate LayerNorm kernel (1.4 ms, 6.4 GB of HBM traffic at B = 4096) disappears.
__global__ void __launch_bounds__(CL_THREADS, 1) conv_ln_taps_persist_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                             const __grid_constant__ CUtensorMap tmB,
                                                                             const __grid_constant__ ConvLnArgs p, int MT) {
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  constexpr int NST = CL_STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NST + 2];
  __shared__ uint32_t tmem_holder;
  __shared__ float part[2][2][128][2];
  __shared__ __align__(16) float s_bias[384], s_gamma[384], s_beta[384];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[NST]);
  const uint32_t bar_acc = smem_u32(&bars[2 * NST]);
  const uint32_t bar_epi = smem_u32(&bars[2 * NST + 1]);
  const uint32_t stg_off = (uint32_t)NST * CL_STAGE_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(bar_epi, 8);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_holder;

  if (warp == 0) {
    const uint32_t a_box_bytes = (uint32_t)TC_BK * 2u * (uint32_t)p.Lbox * (uint32_t)p.Bbox;
    int it = 0;
    for (int mt = blockIdx.x; mt < MT; mt += gridDim.x) {
      const int sample0 = (mt / p.tps) * p.Bbox, l0 = (mt % p.tps) * p.Lbox;
      for (int jj = 0; jj < 3; ++jj) {
        const int j = 2 - jj;
        for (int kb = p.kb_lo[j]; kb < p.kb_hi[j]; ++kb, ++it) {
          const int tap = kb / p.cpt, cc = kb - tap * p.cpt;
          const int s = it % NST;
          const uint32_t ph = (uint32_t)(it / NST) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + s * CL_STAGE_BYTES;
            mbar_expect_tx(bar_full + 8 * s, a_box_bytes + CL_SUB_BYTES);
            tma_load_3d(sa, &tmA, cc * TC_BK, l0 + tap - p.pad, sample0, bar_full + 8 * s);
            tma_load_2d(sa + CL_A_BYTES, &tmB, tap * p.Cin + cc * TC_BK, p.brow_base[j], bar_full + 8 * s);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    int it = 0, n = 0;
    for (int mt = blockIdx.x; mt < MT; mt += gridDim.x, ++n) {
      for (int jj = 0; jj < 3; ++jj) {
        const int j = 2 - jj;
        if (jj == 1 && n > 0) {  // columns [0,256) are shared with the previous tile: wait for its epilogue
          mbar_wait_sleep(bar_epi, ((uint32_t)(n - 1)) & 1u);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + (j == 2 ? 256u + 128u * (uint32_t)(n & 1) : 128u * (uint32_t)j);
        for (int kb = p.kb_lo[j]; kb < p.kb_hi[j]; ++kb, ++it) {
          const int s = it % NST;
          const uint32_t ph = (uint32_t)(it / NST) & 1u;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + s * CL_STAGE_BYTES;
            const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + CL_A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, IDESC, (kb > p.kb_lo[j] || k > 0) ? 1u : 0u);
            umma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
        }
      }
      if (elect_one_sync()) umma_commit(bar_acc);
      __syncwarp();
    }
  } else {
    {
      const int gw0 = 128 / p.ng;
      for (int i = threadIdx.x - 64; i < 384; i += 256) {
        const int j = i >> 7, c = i & 127;
        s_bias[i] = __ldg(p.bias + p.brow_base[j] + c);
        const int g = c / gw0, cc = c - g * gw0;
        s_gamma[i] = __ldg(p.gamma + j * gw0 + cc);
        s_beta[i] = __ldg(p.beta + j * gw0 + cc);
      }
      asm volatile("bar.sync 9, 256;" ::: "memory");
    }
    uint8_t* stg = smem_raw + (smem_base - smem_u32(smem_raw)) + stg_off + (size_t)(warp - 2) * CL_STG_BYTES;
    int n = 0;
    for (int mt = blockIdx.x; mt < MT; mt += gridDim.x, ++n) {
      mbar_wait_sleep(bar_acc, (uint32_t)n & 1u);
      tc_fence_after();
      cl_epilogue_tile<2>(p, mt, 0, tmem_base, 256u + 128u * (uint32_t)(n & 1), s_bias, s_gamma, s_beta, part, stg, warp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) cl_mbar_arrive(bar_epi);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

extern "C" int acb_spectra_conv_ln_bf16(const void* A, const void* Bw, void* out, int nbatch, int L, int Cin, int taps, int pad,
                                        long long a_batch_stride, long long a_row_stride, int ldb, int b_rows, const int* kb_ranges_host,
                                        const int* brow_base_host, int brow_stride_y, int grid_y, int ng, long long row_mul,
                                        long long row_add_y, long long out_rows, const float* bias, const float* gamma,
                                        const float* beta, float eps, const void* down_w, const float* down_bias, void* down_out,
                                        void* stream) {
  ACB_CHECK(A && Bw && (out || down_out) && bias && gamma && beta && kb_ranges_host && brow_base_host, "acb_spectra_conv_ln_bf16: null argument");
  ACB_CHECK((down_w == nullptr) == (down_out == nullptr) && (down_w == nullptr) == (down_bias == nullptr),
            "acb_spectra_conv_ln_bf16: down_w / down_bias / down_out go together");
  ACB_CHECK(nbatch > 0 && L > 0 && Cin > 0 && taps > 0 && (ng == 1 || ng == 2) && grid_y >= 1, "acb_spectra_conv_ln_bf16: bad shape");
  ACB_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bw % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)down_w % 16 == 0) &&
                ((uintptr_t)down_out % 16 == 0) && ((uintptr_t)down_bias % 16 == 0) && a_row_stride % 8 == 0 && a_batch_stride % 8 == 0 && ldb % 8 == 0,
            "acb_spectra_conv_ln_bf16: alignment");
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  ACB_CHECK(enc != nullptr, "acb_spectra_conv_ln_bf16: cuTensorMapEncodeTiled unavailable");
  ConvLnArgs args;
  memset(&args, 0, sizeof(args));
  args.Lbox = L >= TC_BM ? TC_BM : L;
  args.Bbox = L >= TC_BM ? 1 : (TC_BM / L);
  args.tps = L >= TC_BM ? cdiv(L, TC_BM) : 1;
  args.nbatch = nbatch; args.L = L; args.taps = taps; args.pad = pad; args.cpt = cdiv(Cin, TC_BK); args.Cin = Cin;
  const int kb_total = taps * args.cpt;
  for (int j = 0; j < 3; ++j) {
    args.kb_lo[j] = kb_ranges_host[2 * j];
    args.kb_hi[j] = kb_ranges_host[2 * j + 1];
    args.brow_base[j] = brow_base_host[j];
    ACB_CHECK(args.kb_lo[j] >= 0 && args.kb_hi[j] <= kb_total && args.kb_lo[j] < args.kb_hi[j], "acb_spectra_conv_ln_bf16: bad K range %d", j);
  }
  ACB_CHECK(args.kb_lo[2] <= args.kb_lo[0] && args.kb_lo[2] <= args.kb_lo[1] && args.kb_hi[2] >= args.kb_hi[0] && args.kb_hi[2] >= args.kb_hi[1],
            "acb_spectra_conv_ln_bf16: sub-tile 2 must span the other K ranges");
  args.brow_stride_y = brow_stride_y; args.ng = ng; args.row_mul = row_mul; args.row_add_y = row_add_y; args.out_rows = out_rows;
  args.ldc = 3 * (128 / ng); args.bias = bias; args.gamma = gamma; args.beta = beta; args.eps = eps; args.out = (bf16*)out;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)L, (cuuint64_t)nbatch};
    cuuint64_t strides[2] = {(cuuint64_t)a_row_stride * 2, (cuuint64_t)(nbatch > 1 ? a_batch_stride : (long long)a_row_stride * L) * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)args.Lbox, (cuuint32_t)args.Bbox};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACB_CHECK(r == CUDA_SUCCESS, "acb_spectra_conv_ln_bf16: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)((long long)taps * Cin), (cuuint64_t)b_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(Bw), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACB_CHECK(r == CUDA_SUCCESS, "acb_spectra_conv_ln_bf16: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  constexpr size_t smem = (size_t)CL_STAGES * CL_STAGE_BYTES + 8 * CL_STG_BYTES + 4608 + 1024;
  static bool configured = false;
  if (!configured) {
    ACB_CUDA(cudaFuncSetAttribute(conv_ln_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACB_CUDA(cudaFuncSetAttribute(conv_ln_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // polyphase view (rows overlap by construction: row stride 8 samples, 1 tap, Cin = window length): A straight from smem
  const bool hankel = taps == 1 && a_row_stride == 8 && ng == 2 && Cin <= 64 * 17 && L % 128 == 0;
  args.xwin = (const bf16*)A; args.xwin_stride = a_batch_stride; args.xwin_total = (long long)nbatch * a_batch_stride;
  const long long MT = (long long)cdiv(nbatch, args.Bbox) * args.tps;
  ACB_CHECK(MT < (1LL << 31) && grid_y <= 65535, "acb_spectra_conv_ln_bf16: grid too large");
  static int persist = -1;
  if (persist < 0) {
    const char* e = getenv("ACB_CONVLN_PERSIST");
    persist = e ? atoi(e) : 1;
  }
  const bool can_persist = hankel && persist && grid_y <= 4 && MT >= 148 && args.kb_hi[2] - args.kb_lo[2] <= 17;
  if (down_w) {
    // fused 1x1 downsample (3*64 -> 64) + max over the CTA's two phases: only in the persistent stage-0 kernel
    ACB_CHECK(can_persist && grid_y == 4 && ng == 2 && row_mul == 8 && row_add_y == 2,
              "acb_spectra_conv_ln_bf16: the fused downsample needs the persistent polyphase kernel (>= 148 signal windows, 8 phases)");
    args.down_bias = down_bias;
    args.down_out = (bf16*)down_out;
    CUtensorMap tmD;
    cuuint64_t dims[2] = {192, 64};
    cuuint64_t strides[1] = {192 * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(down_w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACB_CHECK(r == CUDA_SUCCESS, "acb_spectra_conv_ln_bf16: cuTensorMapEncodeTiled(down_w) failed with %d", (int)r);
    constexpr size_t dsmem = (size_t)8 * CL_SUB_BYTES + 49152 + 24576 + 2 * CLP_WIN_BYTES + 1024;
    static bool dconf = false;
    if (!dconf) {
      ACB_CUDA(cudaFuncSetAttribute(conv_ln_hankel_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
      dconf = true;
    }
    conv_ln_hankel_persist_kernel<true><<<148, CLP_THREADS, dsmem, (cudaStream_t)stream>>>(tmB, tmD, args, (int)MT, grid_y);
  } else if (can_persist) {
    constexpr size_t psmem = (size_t)CLP_NST * CL_SUB_BYTES + 4 * CLP_NPQ * CL_STG_BYTES + 2 * CLP_WIN_BYTES + 1024;
    static bool pconf = false;
    if (!pconf) {
      ACB_CUDA(cudaFuncSetAttribute(conv_ln_hankel_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      pconf = true;
    }
    conv_ln_hankel_persist_kernel<false><<<148, CLP_THREADS, psmem, (cudaStream_t)stream>>>(tmB, tmB, args, (int)MT, grid_y);
  } else if (hankel) {
    conv_ln_tc_kernel<true><<<dim3((unsigned)MT, (unsigned)grid_y), CL_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, args);
  } else if (persist && grid_y == 1 && MT >= 148) {
    static bool tconf = false;
    if (!tconf) {
      ACB_CUDA(cudaFuncSetAttribute(conv_ln_taps_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tconf = true;
    }
    conv_ln_taps_persist_kernel<<<148, CL_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, args, (int)MT);
  }
  else
    conv_ln_tc_kernel<false><<<dim3((unsigned)MT, (unsigned)grid_y), CL_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, args);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}


extern "C" int acb_debug_timing(int enable, unsigned long long* out4_host) {
  // enable != 0: switch the conv+LN phase timers on/off; out4_host != NULL: copy + reset the accumulators
  ACB_CUDA(cudaDeviceSynchronize());
  ACB_CUDA(cudaMemcpyToSymbol(g_cl_timing_on, &enable, sizeof(int)));
  if (out4_host) {
    ACB_CUDA(cudaMemcpyFromSymbol(out4_host, g_cl_timing, sizeof(unsigned long long) * 4));
    unsigned long long z[4] = {0, 0, 0, 0};
    ACB_CUDA(cudaMemcpyToSymbol(g_cl_timing, z, sizeof(z)));
  }
  return ACB_OK;
}
