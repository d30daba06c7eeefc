// Packed multi-sequence varlen attention on tcgen05 (bf16 in, fp32 softmax, bf16 out), head_dim 16.
//
// ZTF light curves are short (median ~40 tokens, max 258): one CTA per (sequence, head) -- the round-1 kernel -- spends its time
// on CTA set-up and fills 58 of the 128 rows of a UMMA tile.  Here
//   * a PLAN kernel (once per batch, reused by every layer) greedily packs CONSECUTIVE whole sequences into tiles of <= 128
//     packed token rows (queries and keys of a tile are the same rows; the mask is block-diagonal);
//   * one CTA owns (tile, group of 4 heads).  TMA brings the Q / K / V columns of its heads for the 128 rows into shared
//     memory as boxes of {8 elements x 128 rows}: each lands as [128 rows][16 B], a column of UMMA core matrices, so the 8
//     boxes of an operand form [chunk][row][16 B] -- exactly the canonical no-swizzle layout (K-major for Q and K with
//     LBO = 2048 B between the two k-chunks of a head and SBO = 128 B between 8-row groups; MN-major for V, so P V needs no
//     transpose).  No thread touches global memory for the operands.
//   * per head: S = Q K^T is one UMMA (M 128, N = padded rows, K 16) into TMEM; thread r owns query row r: masked max and
//     exp/sum over ITS sequence's key range only (single pass over <= 128 keys, straight from TMEM), P (bf16) goes to shared
//     memory as the K-major A operand, O_h += P V_h by N = 16 UMMAs into its own 16 TMEM columns; the next head's S is issued
//     behind them, so the tensor pipe never waits for the softmax of another head.
//   * the four heads' O sit side by side in TMEM: one 128-byte contiguous store per row at the end.
// Sequences longer than 128 tokens (3 % of a ZTF batch) keep the per-(sequence, head) kernel of attention_tc.cu, driven by the
// plan's list.  Dropout (training) uses the same counter hash as every other attention kernel here, so any backward matches.
#include <stdlib.h>

#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int AP_DH = 16;
constexpr int AP_ROWS = 128;   // packed rows per tile = UMMA M
constexpr int AP_HG = 4;       // heads per CTA
constexpr int AP_THREADS = 256;  // two warpgroups: each owns one half of the key columns of S
constexpr uint32_t AP_OPER_BYTES = AP_HG * 2 * AP_ROWS * 16;  // 16 KB: [8 chunks][128 rows][16 B]
constexpr uint32_t AP_P_BYTES = 16 * AP_ROWS * 16;            // 32 KB: [16 key chunks][128 rows][16 B]
constexpr uint32_t AP_SMEM = 3 * AP_OPER_BYTES + AP_P_BYTES + 1024;

__device__ __forceinline__ unsigned ap_hash(unsigned long long seed, int bh, int i, int j) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)bh << 26) | ((unsigned long long)i << 13) | (unsigned long long)j);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}
// canonical no-swizzle descriptors (cute/atom/mma_traits_sm100.hpp):
//   K-major  ((8,m),(T,2)):((1T,SBO),(1,LBO))      -- 8-row groups SBO apart, the two 16-byte K chunks LBO apart
//   MN-major ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO))  -- 8-element MN blocks SBO apart, 8-row K groups LBO apart
__device__ __forceinline__ uint64_t ap_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t ap_idesc(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---- plan: plan[0] = n_tiles, plan[1] = n_long, plan[2 + 2t] = first sequence of tile t, plan[3 + 2t] = sequences in it,
//            plan[2 + 2*max_tiles + i] = i-th long sequence.
// A tile is a maximal run of consecutive sequences with <= 128 rows in total, each <= 128 rows, at most 128 sequences (greedy,
// left to right).  Greedy packing is a chain (the next tile starts where this one ends), so instead of walking it serially
// (230 us for 4096 sequences) every sequence b computes where a tile starting AT b would end -- nxt[b], a binary search in the
// prefix sums -- and the chain from sequence 0 is marked by pointer doubling in log2(B) parallel rounds. ------------------------
constexpr int AP_PLAN_THREADS = 1024;
constexpr int AP_PLAN_MAX_B = 8192;  // sequences per batch the parallel planner holds in shared memory

__device__ __forceinline__ int ap_tile_end(const int* pre, int B, int b) {
  // largest e in (b, B] with pre[e] - pre[b] <= 128 rows and e - b <= 128 sequences; b itself is known to fit
  int lo = b + 1, hi = min(B, b + AP_ROWS);
  const int lim = pre[b] + AP_ROWS;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (pre[mid] <= lim) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(AP_PLAN_THREADS) attn_plan_kernel(const int* __restrict__ cu, int B, int max_tiles, int* __restrict__ plan) {
  extern __shared__ int sm_plan[];
  int* pre = sm_plan;                 // [B + 1] prefix rows
  int* ja = pre + (B + 1);            // [B + 1] jump pointers (double buffered)
  int* jb = ja + (B + 1);
  int* nxt = jb + (B + 1);            // [B]
  unsigned char* mark = reinterpret_cast<unsigned char*>(nxt + B);  // [B + 1]
  __shared__ int s_warp[2][32];
  __shared__ int s_tot[2];
  const int tid = threadIdx.x;
  for (int b = tid; b <= B; b += AP_PLAN_THREADS) { pre[b] = cu[b]; mark[b] = 0; }
  __syncthreads();
  for (int b = tid; b < B; b += AP_PLAN_THREADS) {
    const int n = pre[b + 1] - pre[b];
    const int e = (n > AP_ROWS) ? b + 1 : ap_tile_end(pre, B, b);  // a long sequence is a chain node of its own
    nxt[b] = e;
    ja[b] = e;
  }
  if (tid == 0) { ja[B] = B; jb[B] = B; mark[0] = 1; }
  __syncthreads();
  int* jc = ja;
  int* jn = jb;
  for (int span = 1; span < B; span <<= 1) {  // after the round with jump length `span`, nodes < 2*span steps from 0 are marked
    for (int b = tid; b < B; b += AP_PLAN_THREADS)
      if (mark[b] && jc[b] < B) mark[jc[b]] = 1;
    for (int b = tid; b < B; b += AP_PLAN_THREADS) jn[b] = jc[jc[b]];
    __syncthreads();
    int* t = jc; jc = jn; jn = t;
  }
  // ---- compact the marked chain nodes: tiles (<= 128 rows) and long sequences, both in sequence order ----
  constexpr int PER = AP_PLAN_MAX_B / AP_PLAN_THREADS;
  int ct = 0, cl = 0;
  const int b0 = tid * PER;
  for (int i = 0; i < PER; ++i) {
    const int b = b0 + i;
    if (b < B && mark[b]) {
      if (pre[b + 1] - pre[b] > AP_ROWS) ++cl;
      else ++ct;
    }
  }
  const int lane = tid & 31, w = tid >> 5;
  int it = ct, il = cl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, it, o), c = __shfl_up_sync(0xffffffffu, il, o);
    if (lane >= o) { it += a; il += c; }
  }
  if (lane == 31) { s_warp[0][w] = it; s_warp[1][w] = il; }
  __syncthreads();
  if (w == 0) {
    int a = s_warp[0][lane], c = s_warp[1][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a2 = __shfl_up_sync(0xffffffffu, a, o), c2 = __shfl_up_sync(0xffffffffu, c, o);
      if (lane >= o) { a += a2; c += c2; }
    }
    s_warp[0][lane] = a; s_warp[1][lane] = c;
    if (lane == 31) { s_tot[0] = a; s_tot[1] = c; }
  }
  __syncthreads();
  int ot = it - ct + (w > 0 ? s_warp[0][w - 1] : 0), ol = il - cl + (w > 0 ? s_warp[1][w - 1] : 0);
  int* tiles = plan + 2;
  int* longs = plan + 2 + 2 * max_tiles;
  for (int i = 0; i < PER; ++i) {
    const int b = b0 + i;
    if (b < B && mark[b]) {
      if (pre[b + 1] - pre[b] > AP_ROWS) longs[ol++] = b;
      else {
        if (ot < max_tiles) { tiles[2 * ot] = b; tiles[2 * ot + 1] = nxt[b] - b; }
        ++ot;
      }
    }
  }
  if (tid == 0) {
    plan[0] = min(s_tot[0], max_tiles);  // cannot exceed the host-side bound; never overrun the table
    plan[1] = s_tot[1];
  }
}

// serial fallback for batches of more than AP_PLAN_MAX_B sequences (same packing rule)
__global__ void attn_plan_serial_kernel(const int* __restrict__ cu, int B, int max_tiles, int* __restrict__ plan) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int nt = 0, nl = 0, first = 0, cnt = 0, rows = 0;
  int* tiles = plan + 2;
  int* longs = plan + 2 + 2 * max_tiles;
  for (int b = 0; b < B; ++b) {
    const int n = cu[b + 1] - cu[b];
    if ((n > AP_ROWS || rows + n > AP_ROWS || cnt >= AP_ROWS) && cnt > 0) {
      if (nt < max_tiles) { tiles[2 * nt] = first; tiles[2 * nt + 1] = cnt; }
      ++nt;
      cnt = 0; rows = 0;
    }
    if (n > AP_ROWS) longs[nl++] = b;
    else {
      if (cnt == 0) first = b;
      ++cnt; rows += n;
    }
  }
  if (cnt > 0) {
    if (nt < max_tiles) { tiles[2 * nt] = first; tiles[2 * nt + 1] = cnt; }
    ++nt;
  }
  plan[0] = min(nt, max_tiles);
  plan[1] = nl;
}

// 256 threads = two warpgroups.  Both see all 128 query rows (TMEM lane = tid & 127); warpgroup g owns the key columns
// [64 g, 64 g + 64) of S: it reduces / exponentiates / writes P for its half, the halves meet through 1 KB of shared memory.
// That doubles the threads working on the softmax (the issue-bound part) without more TMEM or shared memory per CTA, so two
// CTAs still share an SM (2 x 256 TMEM columns, 2 x 81 KB).
__global__ void __launch_bounds__(AP_THREADS, 2) attention_packed_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ cu,
                                                                         const int* __restrict__ plan, int n_heads, float drop_p,
                                                                         AcbSeed seed_s, bf16* __restrict__ out) {
  const int tile = blockIdx.x;
  if (tile >= plan[0]) return;  // uniform: the grid is a host-side upper bound of the tile count
  const unsigned long long seed = seed_s.get();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];  // [0] operands landed, [1] MMA batch retired
  __shared__ uint32_t tmem_holder;
  __shared__ int s_cu[AP_ROWS + 2];
  __shared__ float s_red[2][AP_ROWS];        // per-row partial max of the two key halves
  __shared__ float s_sum[AP_HG][AP_ROWS];    // per-row partial exp-sums of warpgroup 1
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (AP_ROWS - 1), wg = tid >> 7;
  const int hg = blockIdx.y;  // heads hg*4 .. hg*4+3
  const int seq0 = plan[2 + 2 * tile], nseq = plan[3 + 2 * tile];
  const int row0 = cu[seq0];
  const int nrows = cu[seq0 + nseq] - row0;  // <= 128
  const int npad = max(16, (nrows + 15) & ~15);
  const int D = n_heads * AP_DH;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t aQ = base, aK = base + AP_OPER_BYTES, aV = base + 2 * AP_OPER_BYTES, aP = base + 3 * AP_OPER_BYTES;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sV = gen_base + 2 * AP_OPER_BYTES;
  uint8_t* sP = gen_base + 3 * AP_OPER_BYTES;
  const uint32_t bar_load = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i <= nseq && i <= AP_ROWS; i += AP_THREADS) s_cu[i] = cu[seq0 + i] - row0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_holder;          // columns [0, 128): S
  const uint32_t tmem_o = tmem_holder + 128u;   // columns [128, 192): O of the 4 heads
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;

  if (warp == 0 && elect_one_sync()) {
    mbar_expect_tx(bar_load, 3 * AP_OPER_BYTES);
    // one {8 elements x 128 rows} box per 16-byte column chunk: it lands as [128 rows][16 B] = a column of core matrices
#pragma unroll
    for (int op = 0; op < 3; ++op)
#pragma unroll
      for (int c = 0; c < AP_HG * 2; ++c)
        tma_load_2d(base + (uint32_t)op * AP_OPER_BYTES + (uint32_t)c * (AP_ROWS * 16), &tmQKV, op * D + hg * (AP_HG * AP_DH) + c * 8, row0, bar_load);
  }
  // this thread's query row: its sequence and key range (tile-relative)
  int lo = 0, hi = 0, b_seq = seq0;
  if (row < nrows) {
    int j = 0;
    while (j + 1 < nseq && s_cu[j + 1] <= row) ++j;
    lo = s_cu[j];
    hi = s_cu[j + 1];
    b_seq = seq0 + j;
  }
  const unsigned len = (unsigned)(hi - lo);
  const int wlo = __reduce_min_sync(0xffffffffu, row < nrows ? lo : 1 << 30);
  const int whi = __reduce_max_sync(0xffffffffu, row < nrows ? hi : 0);
  // keys that are inside the range of EVERY lane of the warp (empty when the warp straddles sequences or has idle rows):
  // chunks inside need no per-element range predicates
  const int ilo = __reduce_max_sync(0xffffffffu, lo);
  const int ihi = __reduce_min_sync(0xffffffffu, hi);
  const float drop_inv = 1.0f / (1.0f - drop_p);
  const unsigned drop_thr = (unsigned)(drop_p * 4294967296.0);
  const int kbeg = wg * 64, kend = min(npad, kbeg + 64);  // this warpgroup's key columns
  constexpr float SC = 0.25f * 1.4426950408889634f;       // 1/sqrt(dh) * log2(e): p = 2^(s*SC - max*SC)

  mbar_wait(bar_load, 0);
  // V rows in [nrows, npad) belong to the next tile or to unwritten capacity rows: P is zero there, but 0 * NaN is not
  for (int i = tid; i < (npad - nrows) * (AP_HG * 2); i += AP_THREADS) {
    const int r = nrows + i / (AP_HG * 2), c = i % (AP_HG * 2);
    *reinterpret_cast<uint4*>(sV + ((size_t)c * AP_ROWS + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  __syncthreads();
  uint32_t ph = 0;
  if (warp == 0 && elect_one_sync()) {
    tc_fence_after();
    umma_bf16(tmem_s, ap_desc(aQ, AP_ROWS * 16, 128), ap_desc(aK, AP_ROWS * 16, 128), ap_idesc(npad, false), 0u);
    umma_commit(bar_mma);
  }
  float psum[AP_HG];
#pragma unroll
  for (int h = 0; h < AP_HG; ++h) {
    mbar_wait(bar_mma, ph);
    ph ^= 1u;
    tc_fence_after();
    // ---- pass 1: max of row `row` over its own sequence's keys inside this warpgroup's half ----
    float mx = -INFINITY;
    for (int c0 = max(kbeg, wlo & ~31); c0 < min(kend, whi); c0 += 32) {  // warp-uniform: union of the lanes' key ranges
      uint32_t raw[32];
      tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, raw);
      const unsigned off = (unsigned)(c0 - lo);
      if (c0 >= ilo && c0 + 32 <= ihi) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (off + (unsigned)i < len) mx = fmaxf(mx, __uint_as_float(raw[i]));
      }
    }
    s_red[wg][row] = mx;
    __syncthreads();
    const float mxs = fmaxf(s_red[0][row], s_red[1][row]) * SC;
    // ---- pass 2: p = exp(s/4 - max), partial sum, P (bf16) into the K-major A-operand layout ----
    float lsum = 0.0f;
    const int bh = b_seq * n_heads + hg * AP_HG + h;
    const int qi = row - lo;
    for (int c0 = kbeg; c0 < kend; c0 += 32) {
      uint32_t pk[16];
      if (c0 + 32 > wlo && c0 < whi) {
        uint32_t raw[32];
        tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, raw);
        const unsigned off = (unsigned)(c0 - lo);
        const bool interior = c0 >= ilo && c0 + 32 <= ihi;  // warp-uniform
        float2 ls2 = make_float2(0.0f, 0.0f);  // packed fp32: one FFMA2 / FADD2 per pair of scores
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 e2 = __ffma2_rn(make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), make_float2(SC, SC), make_float2(-mxs, -mxs));
          float p0, p1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e2.x));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e2.y));
          if (!interior) {
            p0 = off + (unsigned)i < len ? p0 : 0.0f;
            p1 = off + (unsigned)(i + 1) < len ? p1 : 0.0f;
          }
          ls2 = __fadd2_rn(ls2, make_float2(p0, p1));
          if (drop_p > 0.0f) {
            p0 = ap_hash(seed, bh, qi, (int)off + i) >= drop_thr ? p0 * drop_inv : 0.0f;
            p1 = ap_hash(seed, bh, qi, (int)off + i + 1) >= drop_thr ? p1 * drop_inv : 0.0f;
          }
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
        }
        lsum += ls2.x + ls2.y;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)  // 8 keys per 16-byte chunk: chunk (c0/8 + c), row `row`
        if (c0 + 8 * c < kend)
          *reinterpret_cast<uint4*>(sP + (((c0 >> 3) + c) * AP_ROWS + row) * 16) = make_uint4(pk[c * 4 + 0], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
    }
    psum[h] = lsum;
    if (wg == 1) s_sum[h][row] = lsum;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();  // P complete, every thread is done with S and with s_red
    if (warp == 0 && elect_one_sync()) {
      tc_fence_after();
      const uint32_t vh = aV + (uint32_t)h * (2 * AP_ROWS * 16);
      for (int ks = 0; ks < npad; ks += 16)  // O_h += P[:, 16 keys] V_h[16 keys, :]
        umma_bf16(tmem_o + (uint32_t)(h * AP_DH), ap_desc(aP + (uint32_t)(ks >> 3) * (AP_ROWS * 16), AP_ROWS * 16, 128),
                  ap_desc(vh + (uint32_t)ks * 16, 128, AP_ROWS * 16), ap_idesc(AP_DH, true), ks > 0 ? 1u : 0u);
      if (h + 1 < AP_HG) {
        const uint32_t off = (uint32_t)(h + 1) * (2 * AP_ROWS * 16);
        umma_bf16(tmem_s, ap_desc(aQ + off, AP_ROWS * 16, 128), ap_desc(aK + off, AP_ROWS * 16, 128), ap_idesc(npad, false), 0u);
      }
      umma_commit(bar_mma);
    }
  }
  mbar_wait(bar_mma, ph);
  tc_fence_after();
  // ---- O / sum -> bf16: warpgroup g stores heads 2g, 2g+1 (64 contiguous bytes per row) ----
  if (wg == 0) {
#pragma unroll
    for (int h = 0; h < AP_HG; ++h) s_sum[h][row] += psum[h];  // this row's entries were written by the other warpgroup only
  }
  __syncthreads();
  {
    uint32_t o[32];
    tmem_ld32(tmem_o + lane_addr + (uint32_t)(wg * 32), o);
    if (row < nrows) {
      uint4* dst = reinterpret_cast<uint4*>(out + (long long)(row0 + row) * D + hg * (AP_HG * AP_DH) + wg * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float tot = s_sum[wg * 2 + (q >> 1)][row];
        const float s = tot > 0.0f ? 1.0f / tot : 0.0f;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[q * 8 + 2 * e]) * s, __uint_as_float(o[q * 8 + 2 * e + 1]) * s);
          w[e] = *reinterpret_cast<uint32_t*>(&hh);
        }
        dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(256u) : "memory");
  }
}

// ======================================================================================================================
// Backward of the packed attention on tcgen05.  Same tiles, same plan, same canonical TMA layouts; per (tile, 4 heads) CTA:
//   S  = Q K^T            dP = dO V^T                      (two K = 16 UMMAs into TMEM, recomputed: nothing was saved)
//   threads: row r, 32-key slice: P = softmax(S/4) recomputed (max and sum meet across the 4 slices through shared memory),
//            D = sum_j P_j dP_j, dS = P (dP - D) / 4; P (with the forward's dropout mask) and dS go to shared memory as bf16
//   dV = P^T dO           dQ = dS K           dK = dS^T Q  (N = 16 UMMAs; P^T / dS^T are the SAME buffers read MN-major)
// 512 threads: 16 warps = 4 TMEM lane quarters x 4 key slices, so a thread keeps its 32 scores and 32 dP values in registers
// across the three softmax passes (one TMEM read each).  TMEM: S | dP | dQ[4 heads] | dK | dV = 448 columns.
constexpr int AB_THREADS = 512;
constexpr uint32_t AB_SMEM = 4 * AP_OPER_BYTES + 2 * AP_P_BYTES + 1024;

__device__ __forceinline__ uint32_t ab_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(AB_THREADS, 1) attention_packed_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                             const __grid_constant__ CUtensorMap tmDO, const int* __restrict__ cu,
                                                                             const int* __restrict__ plan, int n_heads, float drop_p,
                                                                             AcbSeed seed_s, bf16* __restrict__ dqkv) {
  const int tile = blockIdx.x;
  if (tile >= plan[0]) return;
  const unsigned long long seed = seed_s.get();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_holder;
  __shared__ int s_cu[AP_ROWS + 2];
  __shared__ float s_red[3][4][AP_ROWS];  // per-row partial max / sum / sum(p dP) of the four key slices
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, sl = warp >> 2;   // TMEM lane quarter, key slice
  const int row = q * 32 + lane;
  const int hg = blockIdx.y;
  const int seq0 = plan[2 + 2 * tile], nseq = plan[3 + 2 * tile];
  const int row0 = cu[seq0];
  const int nrows = cu[seq0 + nseq] - row0;
  const int npad = max(16, (nrows + 15) & ~15);
  const int D = n_heads * AP_DH;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t aQ = base, aK = base + AP_OPER_BYTES, aV = base + 2 * AP_OPER_BYTES, aO = base + 3 * AP_OPER_BYTES;
  const uint32_t aP = base + 4 * AP_OPER_BYTES, aS = aP + AP_P_BYTES;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sP = gen + 4 * AP_OPER_BYTES;
  uint8_t* sS = sP + AP_P_BYTES;
  const uint32_t bar_load = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i <= nseq && i <= AP_ROWS; i += AB_THREADS) s_cu[i] = cu[seq0 + i] - row0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_holder;
  const uint32_t tm_s = tm, tm_dp = tm + 128u, tm_dq = tm + 256u, tm_dk = tm + 320u, tm_dv = tm + 384u;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

  if (warp == 0 && elect_one_sync()) {
    mbar_expect_tx(bar_load, 4 * AP_OPER_BYTES);
#pragma unroll
    for (int op = 0; op < 3; ++op)
#pragma unroll
      for (int c = 0; c < AP_HG * 2; ++c)
        tma_load_2d(base + (uint32_t)op * AP_OPER_BYTES + (uint32_t)c * (AP_ROWS * 16), &tmQKV, op * D + hg * (AP_HG * AP_DH) + c * 8, row0, bar_load);
#pragma unroll
    for (int c = 0; c < AP_HG * 2; ++c) tma_load_2d(aO + (uint32_t)c * (AP_ROWS * 16), &tmDO, hg * (AP_HG * AP_DH) + c * 8, row0, bar_load);
  }
  int lo = 0, hi = 0, b_seq = seq0;
  if (row < nrows) {
    int j = 0;
    while (j + 1 < nseq && s_cu[j + 1] <= row) ++j;
    lo = s_cu[j];
    hi = s_cu[j + 1];
    b_seq = seq0 + j;
  }
  const unsigned len = (unsigned)(hi - lo);
  const int c0 = sl * 32;                       // this thread's key columns [c0, c0 + 32)
  const bool slice_live = c0 < npad;            // warp-uniform
  const unsigned off = (unsigned)(c0 - lo);
  const float drop_inv = 1.0f / (1.0f - drop_p);
  const unsigned drop_thr = (unsigned)(drop_p * 4294967296.0);
  constexpr float SC = 0.25f * 1.4426950408889634f;

  mbar_wait(bar_load, 0);
  // rows in [nrows, npad) belong to the next tile or to capacity rows: keep them out of the MMAs (0 * NaN)
  for (int i = tid; i < (npad - nrows) * (4 * AP_HG * 2); i += AB_THREADS) {
    const int r = nrows + i / (4 * AP_HG * 2), c = i % (4 * AP_HG * 2);  // c over the 32 chunk columns of Q | K | V | dO
    *reinterpret_cast<uint4*>(gen + ((size_t)c * AP_ROWS + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  __syncthreads();
  uint32_t ph = 0;
  if (warp == 0 && elect_one_sync()) {
    tc_fence_after();
    umma_bf16(tm_s, ap_desc(aQ, AP_ROWS * 16, 128), ap_desc(aK, AP_ROWS * 16, 128), ab_idesc(npad, false, false), 0u);
    umma_bf16(tm_dp, ap_desc(aO, AP_ROWS * 16, 128), ap_desc(aV, AP_ROWS * 16, 128), ab_idesc(npad, false, false), 0u);
    umma_commit(bar_mma);
  }
#pragma unroll 1
  for (int h = 0; h < AP_HG; ++h) {
    mbar_wait(bar_mma, ph);
    ph ^= 1u;
    tc_fence_after();
    float sv[32], dv[32];
    float mx = -INFINITY;
    if (slice_live) {
      uint32_t raw[32];
      tmem_ld32(tm_s + lane_addr + (uint32_t)c0, raw);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        sv[i] = __uint_as_float(raw[i]);
        if (off + (unsigned)i < len) mx = fmaxf(mx, sv[i]);
      }
      tmem_ld32(tm_dp + lane_addr + (uint32_t)c0, raw);
#pragma unroll
      for (int i = 0; i < 32; ++i) dv[i] = __uint_as_float(raw[i]);
    }
    s_red[0][sl][row] = mx;
    __syncthreads();
    const float mxs = fmaxf(fmaxf(s_red[0][0][row], s_red[0][1][row]), fmaxf(s_red[0][2][row], s_red[0][3][row])) * SC;
    const int bh = b_seq * n_heads + hg * AP_HG + h;
    const int qi = row - lo;
    float l_part = 0.0f, d_part = 0.0f;
    if (slice_live) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(sv[i], SC, -mxs)));
        p = off + (unsigned)i < len ? p : 0.0f;
        l_part += p;
        float g = dv[i];
        if (drop_p > 0.0f) {  // forward: P_drop = P * keep / (1 - p)  ->  dP = dP_drop * keep / (1 - p)
          const bool keep = ap_hash(seed, bh, qi, (int)off + i) >= drop_thr;
          g = keep ? g * drop_inv : 0.0f;
          dv[i] = g;
        }
        d_part = fmaf(p, g, d_part);
        sv[i] = p;  // unnormalised probability
      }
    }
    s_red[1][sl][row] = l_part;
    s_red[2][sl][row] = d_part;
    __syncthreads();
    const float lsum = (s_red[1][0][row] + s_red[1][1][row]) + (s_red[1][2][row] + s_red[1][3][row]);
    const float dsum = (s_red[2][0][row] + s_red[2][1][row]) + (s_red[2][2][row] + s_red[2][3][row]);
    const float inv_l = lsum > 0.0f ? 1.0f / lsum : 0.0f;
    const float Dr = dsum * inv_l;
    if (slice_live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c0 + 8 * c < npad) {
          uint32_t pw[4], sw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = c * 8 + 2 * e;
            const float p0 = sv[i] * inv_l, p1 = sv[i + 1] * inv_l;
            const float s0 = p0 * (dv[i] - Dr) * 0.25f, s1 = p1 * (dv[i + 1] - Dr) * 0.25f;
            float pd0 = p0, pd1 = p1;
            if (drop_p > 0.0f) {
              pd0 = ap_hash(seed, bh, qi, (int)off + i) >= drop_thr ? p0 * drop_inv : 0.0f;
              pd1 = ap_hash(seed, bh, qi, (int)off + i + 1) >= drop_thr ? p1 * drop_inv : 0.0f;
            }
            __nv_bfloat162 hp = __floats2bfloat162_rn(pd0, pd1), hs = __floats2bfloat162_rn(s0, s1);
            pw[e] = *reinterpret_cast<uint32_t*>(&hp);
            sw[e] = *reinterpret_cast<uint32_t*>(&hs);
          }
          const size_t o = ((size_t)((c0 >> 3) + c) * AP_ROWS + row) * 16;
          *reinterpret_cast<uint4*>(sP + o) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
          *reinterpret_cast<uint4*>(sS + o) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one_sync()) {
      tc_fence_after();
      const uint32_t ho = (uint32_t)h * (2 * AP_ROWS * 16);  // this head's two 16-byte chunks inside an operand
      for (int ks = 0; ks < npad; ks += 16) {
        const uint32_t acc = ks > 0 ? 1u : 0u;
        // dV[key, d] += sum_q P[q, key] dO[q, d]   A = P^T (MN-major view, K = queries), B = dO (MN-major)
        umma_bf16(tm_dv + (uint32_t)(h * AP_DH), ap_desc(aP + (uint32_t)ks * 16, 128, AP_ROWS * 16), ap_desc(aO + ho + (uint32_t)ks * 16, 128, AP_ROWS * 16),
                  ab_idesc(AP_DH, true, true), acc);
        // dQ[q, d] += sum_key dS[q, key] K[key, d]  A = dS (K-major, K = keys), B = K (MN-major)
        umma_bf16(tm_dq + (uint32_t)(h * AP_DH), ap_desc(aS + (uint32_t)(ks >> 3) * (AP_ROWS * 16), AP_ROWS * 16, 128),
                  ap_desc(aK + ho + (uint32_t)ks * 16, 128, AP_ROWS * 16), ab_idesc(AP_DH, false, true), acc);
        // dK[key, d] += sum_q dS[q, key] Q[q, d]    A = dS^T (MN-major view), B = Q (MN-major)
        umma_bf16(tm_dk + (uint32_t)(h * AP_DH), ap_desc(aS + (uint32_t)ks * 16, 128, AP_ROWS * 16), ap_desc(aQ + ho + (uint32_t)ks * 16, 128, AP_ROWS * 16),
                  ab_idesc(AP_DH, true, true), acc);
      }
      if (h + 1 < AP_HG) {
        const uint32_t no = (uint32_t)(h + 1) * (2 * AP_ROWS * 16);
        umma_bf16(tm_s, ap_desc(aQ + no, AP_ROWS * 16, 128), ap_desc(aK + no, AP_ROWS * 16, 128), ab_idesc(npad, false, false), 0u);
        umma_bf16(tm_dp, ap_desc(aO + no, AP_ROWS * 16, 128), ap_desc(aV + no, AP_ROWS * 16, 128), ab_idesc(npad, false, false), 0u);
      }
      umma_commit(bar_mma);
    }
  }
  mbar_wait(bar_mma, ph);
  tc_fence_after();
  // ---- dQ | dK | dV (192 contiguous TMEM columns) -> bf16 rows of dqkv: slice sl stores the 16-column groups 3 sl .. 3 sl + 2 ----
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int g = sl * 3 + t;  // 0..11: array g / 4 (dQ, dK, dV), head g % 4
    uint32_t o[16];
    {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]), "=r"(o[10]),
            "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
          : "r"(tm_dq + lane_addr + (uint32_t)(g * 16))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    if (row < nrows) {
      uint32_t w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1]));
        w[e] = *reinterpret_cast<uint32_t*>(&hh);
      }
      uint4* dst = reinterpret_cast<uint4*>(dqkv + (long long)(row0 + row) * 3 * D + (g >> 2) * D + hg * (AP_HG * AP_DH) + (g & 3) * AP_DH);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

}  // namespace

int acb_attention_tc_long(const void* qkv, const int* cu_seqlens, const int* long_list, const int* n_long_dev, int grid_x, int n_heads,
                          int max_seqlen, float drop_p, long long seed, void* out, cudaStream_t st);  // attention_tc.cu
int acb_attention_bwd_long(const void* qkv, const void* dout, const int* cu_seqlens, const int* long_list, const int* n_long_dev, int grid_x,
                           int n_heads, int max_seqlen, float drop_p, long long seed, void* dqkv, cudaStream_t st);  // backward.cu

extern "C" {

int acb_attention_plan(const int* cu_seqlens, int B, int max_tiles, int* plan, void* stream) {
  ACB_CHECK(cu_seqlens && plan && B > 0 && max_tiles > 0, "acb_attention_plan: bad arguments");
  if (B > AP_PLAN_MAX_B) {
    attn_plan_serial_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cu_seqlens, B, max_tiles, plan);
  } else {
    const size_t smem = (size_t)(4 * (B + 1)) * sizeof(int) + (size_t)(B + 1) + 16;
    ACB_CUDA(cudaFuncSetAttribute(attn_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_plan_kernel<<<1, AP_PLAN_THREADS, smem, (cudaStream_t)stream>>>(cu_seqlens, B, max_tiles, plan);
  }
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_attention_packed(const void* qkv, const int* cu_seqlens, const int* plan, int B, int max_tiles, long long total_rows, int n_heads,
                         int dh, int max_seqlen, float drop_p, long long seed, void* out, void* stream) {
  ACB_CHECK(qkv && cu_seqlens && plan && out && B > 0 && max_tiles > 0 && total_rows > 0, "acb_attention_packed: bad arguments");
  ACB_CHECK(dh == AP_DH && n_heads % AP_HG == 0, "acb_attention_packed: needs head_dim 16 and a head count that is a multiple of 4");
  ACB_CHECK(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), "acb_attention_packed: alignment");
  ACB_CHECK(drop_p >= 0.0f && drop_p < 1.0f, "acb_attention_packed: bad dropout");
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  ACB_CHECK(enc != nullptr, "acb_attention_packed: cuTensorMapEncodeTiled unavailable");
  const int D = n_heads * dh;
  CUtensorMap tm;
  {
    // qkv[T, 3D] row-major; box = {8 elements (one 16-byte chunk), 128 rows}; rows past T are zero-filled
    cuuint64_t dims[2] = {(cuuint64_t)(3 * D), (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)(3 * D) * 2};
    cuuint32_t box[2] = {8, (cuuint32_t)AP_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ACB_CHECK(r == CUDA_SUCCESS, "acb_attention_packed: cuTensorMapEncodeTiled failed with %d", (int)r);
  }
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaFuncSetAttribute(attention_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AP_SMEM));
  static int dbg = -1;  // ACB_ATTN_DEBUG: 1 = packed tiles only, 2 = long sequences only (timing probes; results incomplete)
  if (dbg < 0) { const char* e = getenv("ACB_ATTN_DEBUG"); dbg = e ? atoi(e) : 0; }
  // Sequences longer than one tile run the per-(sequence, head) kernel over the plan's list (grid = host-side upper bound).
  // Both kernels are latency-chain bound and touch disjoint rows, so the long list is FORKED onto a side stream (event
  // fork / join, capturable in a CUDA graph) and the two kernels share the SMs instead of running back to back.
  const int max_long = (int)std::min<long long>((long long)B, total_rows / (AP_ROWS + 1));
  const bool has_long = max_long > 0 && max_seqlen > AP_ROWS && dbg != 1;
  int dev = 0;
  ACB_CUDA(cudaGetDevice(&dev));
  ACB_CHECK(dev >= 0 && dev < 64, "acb_attention_packed: device index %d out of range", dev);
  static cudaStream_t side[64];
  static cudaEvent_t ev_fork[64], ev_join[64];
  if (has_long) {
    if (!side[dev]) {
      ACB_CUDA(cudaStreamCreateWithFlags(&side[dev], cudaStreamNonBlocking));
      ACB_CUDA(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
      ACB_CUDA(cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming));
    }
    ACB_CUDA(cudaEventRecord(ev_fork[dev], st));
    ACB_CUDA(cudaStreamWaitEvent(side[dev], ev_fork[dev], 0));
    const int rc = acb_attention_tc_long(qkv, cu_seqlens, plan + 2 + 2 * max_tiles, plan + 1, max_long, n_heads, max_seqlen, drop_p, seed, out,
                                         side[dev]);
    if (rc != ACB_OK) return rc;
    ACB_CUDA(cudaEventRecord(ev_join[dev], side[dev]));
  }
  if (dbg != 2) {
    attention_packed_kernel<<<dim3(max_tiles, n_heads / AP_HG), AP_THREADS, AP_SMEM, st>>>(tm, cu_seqlens, plan, n_heads, drop_p, acb_seed(seed),
                                                                                        (bf16*)out);
    ACB_LAUNCH_CHECK();
    acb_count_launch();
  }
  if (has_long) ACB_CUDA(cudaStreamWaitEvent(st, ev_join[dev], 0));
  return ACB_OK;
}

int acb_attention_packed_bwd(const void* qkv, const void* dout, const int* cu_seqlens, const int* plan, int B, int max_tiles,
                             long long total_rows, int n_heads, int dh, int max_seqlen, float drop_p, long long seed, void* dqkv, void* stream) {
  ACB_CHECK(qkv && dout && cu_seqlens && plan && dqkv && B > 0 && max_tiles > 0 && total_rows > 0, "acb_attention_packed_bwd: bad arguments");
  ACB_CHECK(dh == AP_DH && n_heads % AP_HG == 0, "acb_attention_packed_bwd: needs head_dim 16 and a head count that is a multiple of 4");
  ACB_CHECK((((uintptr_t)qkv | (uintptr_t)dout | (uintptr_t)dqkv) % 16) == 0, "acb_attention_packed_bwd: alignment");
  ACB_CHECK(drop_p >= 0.0f && drop_p < 1.0f, "acb_attention_packed_bwd: bad dropout");
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  ACB_CHECK(enc != nullptr, "acb_attention_packed_bwd: cuTensorMapEncodeTiled unavailable");
  const int D = n_heads * dh;
  CUtensorMap tmQ, tmO;
  auto mk = [&](CUtensorMap* tm, const void* ptr, int cols) -> CUresult {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)total_rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {8, (cuuint32_t)AP_ROWS};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  ACB_CHECK(mk(&tmQ, qkv, 3 * D) == CUDA_SUCCESS && mk(&tmO, dout, D) == CUDA_SUCCESS, "acb_attention_packed_bwd: cuTensorMapEncodeTiled failed");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaFuncSetAttribute(attention_packed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AB_SMEM));
  attention_packed_bwd_kernel<<<dim3(max_tiles, n_heads / AP_HG), AB_THREADS, AB_SMEM, st>>>(tmQ, tmO, cu_seqlens, plan, n_heads, drop_p, acb_seed(seed),
                                                                                         (bf16*)dqkv);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  const int max_long = (int)std::min<long long>((long long)B, total_rows / (AP_ROWS + 1));
  if (max_long > 0 && max_seqlen > AP_ROWS)
    return acb_attention_bwd_long(qkv, dout, cu_seqlens, plan + 2 + 2 * max_tiles, plan + 1, max_long, n_heads, max_seqlen, drop_p, seed, dqkv, st);
  return ACB_OK;
}

}  // extern "C"
