// Fused masked varlen attention on tcgen05 (bf16 in, fp32 softmax, bf16 out) for head_dim 16.
//
// One CTA per (sequence, head), 4 warps = 128 query rows per tile.  Per query tile:
//   Q (pre-scaled by 1/sqrt(dh)), K, V^T of the head are laid out in shared memory as no-swizzle K-major UMMA
//   core matrices (16-byte row pieces 16 B apart, SBO = 128 B);
//   S = Q K^T is ONE tcgen05.mma (K = 16 = head_dim, N = padded key count) into TMEM;
//   every thread owns one query row: max and sum straight from TMEM (exact two-pass softmax, fp32),
//   P (bf16) goes back to shared memory in 64-key blocks (double buffered) and each block is multiplied with V by
//   four N = 16 tcgen05.mma accumulating O in TMEM; O / sum is written as bf16.
// Padding never exists here: sequences are varlen-packed (cu_seqlens), keys beyond the length are zero rows that
// the softmax skips.  Training-time dropout on P uses the same counter hash as the CUDA-core kernels.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int AT_DH = 16;
constexpr int AT_THREADS = 128;
constexpr int AT_CH = 96;  // keys per S chunk

__device__ __forceinline__ unsigned attn_hash_tc(unsigned long long seed, int bh, int i, int j) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)bh << 26) | ((unsigned long long)i << 13) | (unsigned long long)j);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}

// K-major, no swizzle: start | LBO (between the two 16-byte K chunks of one MMA) | SBO (between 8-row groups)
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared memory (dynamic): Q [2][128] x 16 B | K [2][ncap] x 16 B | Vt [ncap/8][16] x 16 B | P [2 buffers][8][128] x 16 B
__global__ void __launch_bounds__(AT_THREADS) attention_tc_kernel(const bf16* __restrict__ qkv, const int* __restrict__ cu, int n_heads,
                                                                  int ncap, float drop_p, AcbSeed seed_s, bf16* __restrict__ out,
                                                                  const int* __restrict__ seq_list, const int* __restrict__ n_list) {
  // seq_list (optional): blockIdx.x indexes a device-side list of sequences (the long sequences of attention_packed.cu's plan);
  // the grid is then a host-side upper bound and the blocks past *n_list leave at once
  if (seq_list && (int)blockIdx.x >= *n_list) return;
  const unsigned long long seed = seed_s.get();
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bars[3];  // [0] S ready, [1..2] P buffer consumed
  __shared__ uint32_t tmem_holder;
  const int b = seq_list ? seq_list[blockIdx.x] : (int)blockIdx.x;
  const int t0 = cu[b], n = cu[b + 1] - t0;
  const int D = n_heads * AT_DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npad = (n + 15) & ~15;
  uint8_t* sQ = sm;
  uint8_t* sK = sQ + 2 * 128 * 16;
  uint8_t* sV = sK + 2 * ncap * 16;
  uint8_t* sP = sV + (ncap / 8) * 256;
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
  const uint32_t bar_s = smem_u32(&bars[0]), bar_p = smem_u32(&bars[1]);

  // S is produced in chunks of <= AT_CH keys (columns [0, 96)), O lives in columns [96, 112): 128 TMEM columns for EVERY
  // sequence length, so four CTAs share an SM.  (One S tile over all keys made the 3 % longest sequences allocate all 512
  // columns and run alone on their SM -- half of the kernel's time.)  Longer sequences recompute S in the second pass:
  // the MMA is free, TMEM occupancy is not.
  const uint32_t ncols = 128;
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 1);
    mbar_init(bar_p + 8, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_holder;
  const uint32_t tmem_o = tmem_s + (uint32_t)AT_CH;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;

  const float drop_inv = 1.0f / (1.0f - drop_p);
  const unsigned drop_thr = (unsigned)(drop_p * 4294967296.0);
  uint32_t ph_s = 0, ph_p[2] = {0, 0};
  int p_uses[2] = {0, 0};  // outstanding commit per P buffer

  {
    const int h = blockIdx.y;  // one (sequence, head) per CTA: 8x finer granularity balances the lognormal lengths
    // ---- K and V^T of this head (all keys) ----
    for (int i = tid; i < npad * 2; i += AT_THREADS) {  // K: [chunk c][key] 16-byte pieces
      const int key = i >> 1, c = i & 1;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (key < n) v = *reinterpret_cast<const uint4*>(qkv + (long long)(t0 + key) * 3 * D + D + h * AT_DH + c * 8);
      *reinterpret_cast<uint4*>(sK + (c * ncap + key) * 16) = v;
    }
    for (int i = tid; i < npad * 2; i += AT_THREADS) {  // V^T: element (d, key) at [key/8][d] 16-byte row, position key%8
      const int key = i >> 1, c = i & 1;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (key < n) v = *reinterpret_cast<const uint4*>(qkv + (long long)(t0 + key) * 3 * D + 2 * D + h * AT_DH + c * 8);
      const uint16_t* e = reinterpret_cast<const uint16_t*>(&v);
      uint16_t* dst = reinterpret_cast<uint16_t*>(sV + (key >> 3) * 256) + (key & 7);
#pragma unroll
      for (int d = 0; d < 8; ++d) dst[(c * 8 + d) * 8] = e[d];
    }
    for (int q0 = 0; q0 < n; q0 += 128) {
      // ---- Q tile, pre-scaled by 1/sqrt(16) = 0.25 (exact in bf16) ----
      for (int i = tid; i < 256; i += AT_THREADS) {
        const int r = i >> 1, c = i & 1;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q0 + r < n) {
          v = *reinterpret_cast<const uint4*>(qkv + (long long)(t0 + q0 + r) * 3 * D + h * AT_DH + c * 8);
          __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
          for (int k = 0; k < 4; ++k) hp[k] = __hmul2(hp[k], __floats2bfloat162_rn(0.25f, 0.25f));
        }
        *reinterpret_cast<uint4*>(sQ + (c * 128 + r) * 16) = v;
      }
      fence_proxy_async();
      __syncthreads();
      // ---- pass 1: row maxima over all keys, S = Q K^T one chunk (one UMMA, K = 16) at a time ----
      const bool single = npad <= AT_CH;
      const int qi = q0 + tid;
      float mx = -INFINITY;
      for (int kc = 0; kc < npad; kc += AT_CH) {
        const int nn = min(AT_CH, npad - kc);
        if (kc > 0) {  // every thread has read the previous chunk
          tc_fence_before();
          __syncthreads();
        }
        if (warp == 0 && elect_one_sync()) {
          tc_fence_after();
          umma_bf16(tmem_s, desc_nosw(aQ, 128 * 16, 128), desc_nosw(aK + kc * 16, ncap * 16, 128), idesc_bf16(nn), 0u);
          umma_commit(bar_s);
        }
        mbar_wait(bar_s, ph_s);
        ph_s ^= 1u;
        tc_fence_after();
        for (int c0 = 0; c0 < nn; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_s + lane_addr + (uint32_t)c0, raw);
          if (kc + c0 + 32 <= n && c0 + 32 <= nn) {  // interior chunk (every key real): no per-element predicates
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kc + c0 + i < n && c0 + i < nn) mx = fmaxf(mx, __uint_as_float(raw[i]));
          }
        }
      }
      // ---- pass 2: P = exp(S - max) in 64-key blocks -> smem -> O += P V ----
      float lsum = 0.0f;
      const int bh = b * n_heads + h;
      int blk = 0;
      for (int kc = 0; kc < npad; kc += AT_CH) {
        const int nn = min(AT_CH, npad - kc);
        if (!single) {  // S of this chunk again (the single-chunk case still holds it)
          tc_fence_before();
          __syncthreads();
          if (warp == 0 && elect_one_sync()) {
            tc_fence_after();
            umma_bf16(tmem_s, desc_nosw(aQ, 128 * 16, 128), desc_nosw(aK + kc * 16, ncap * 16, 128), idesc_bf16(nn), 0u);
            umma_commit(bar_s);
          }
          mbar_wait(bar_s, ph_s);
          ph_s ^= 1u;
          tc_fence_after();
        }
        for (int kl = 0; kl < nn; kl += 64, ++blk) {
          const int k0 = kc + kl;
          const int buf = blk & 1;
          if (p_uses[buf]) {  // the MMAs that read this P buffer must have retired
            mbar_wait(bar_p + 8 * buf, ph_p[buf]);
            ph_p[buf] ^= 1u;
            p_uses[buf] = 0;
          }
          uint8_t* pb = sP + buf * (8 * 128 * 16);
          const int kend = min(64, nn - kl);
          for (int c0 = 0; c0 < kend; c0 += 32) {
            uint32_t raw[32];
            tmem_ld32(tmem_s + lane_addr + (uint32_t)(kl + c0), raw);
            uint32_t pk[16];
            const bool interior = k0 + c0 + 32 <= n && kl + c0 + 32 <= nn;  // warp-uniform: every key of the chunk is real
            constexpr float LOG2E = 1.4426950408889634f;
            const float mx2 = mx * LOG2E;
            if (interior) {  // no per-element range predicates (all chunks of a long sequence but its last one)
#pragma unroll
              float2 ls2 = make_float2(0.0f, 0.0f);  // packed fp32: one FFMA2 / FADD2 per pair of scores
              for (int i = 0; i < 32; i += 2) {
                const int j0 = k0 + c0 + i;
                const float2 e2 = __ffma2_rn(make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), make_float2(LOG2E, LOG2E),
                                             make_float2(-mx2, -mx2));
                float p0, p1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(e2.x));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(e2.y));
                ls2 = __fadd2_rn(ls2, make_float2(p0, p1));
                if (drop_p > 0.0f) {
                  p0 = attn_hash_tc(seed, bh, qi, j0) >= drop_thr ? p0 * drop_inv : 0.0f;
                  p1 = attn_hash_tc(seed, bh, qi, j0 + 1) >= drop_thr ? p1 * drop_inv : 0.0f;
                }
                __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
              }
              lsum += ls2.x + ls2.y;
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const int j0 = k0 + c0 + i;
                const bool in0 = j0 < n && kl + c0 + i < nn, in1 = j0 + 1 < n && kl + c0 + i + 1 < nn;  // inside the sequence AND this chunk
                float p0 = in0 ? __expf(__uint_as_float(raw[i]) - mx) : 0.0f;
                float p1 = in1 ? __expf(__uint_as_float(raw[i + 1]) - mx) : 0.0f;
                lsum += p0 + p1;
                if (drop_p > 0.0f) {
                  p0 = attn_hash_tc(seed, bh, qi, j0) >= drop_thr ? p0 * drop_inv : 0.0f;
                  p1 = attn_hash_tc(seed, bh, qi, j0 + 1) >= drop_thr ? p1 * drop_inv : 0.0f;
                }
                __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
              }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)  // 8 keys per 16-byte chunk: chunk (c0/8 + c), row tid
              *reinterpret_cast<uint4*>(pb + (((c0 >> 3) + c) * 128 + tid) * 16) = make_uint4(pk[c * 4 + 0], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
          }
          fence_proxy_async();
          tc_fence_before();
          __syncthreads();
          if (warp == 0 && elect_one_sync()) {
            tc_fence_after();
            for (int kk = 0; kk < kend; kk += 16) {  // O += P[:, 16 keys] V[16 keys, :]
              const uint32_t chunk = (uint32_t)(kk >> 3);
              umma_bf16(tmem_o, desc_nosw(aP + buf * (8 * 128 * 16) + chunk * (128 * 16), 128 * 16, 128),
                        desc_nosw(aV + (uint32_t)((k0 + kk) >> 3) * 256, 256, 128), idesc_bf16(16), (k0 + kk) > 0 ? 1u : 0u);
            }
            umma_commit(bar_p + 8 * buf);
          }
          p_uses[buf] = 1;
        }
      }
      // ---- drain, normalise, store ----
      for (int buf = 0; buf < 2; ++buf) {
        if (p_uses[buf]) {
          mbar_wait(bar_p + 8 * buf, ph_p[buf]);
          ph_p[buf] ^= 1u;
          p_uses[buf] = 0;
        }
      }
      tc_fence_after();
      uint32_t o[16];
      tmem_ld16(tmem_o + lane_addr, o);
      if (qi < n) {
        const float inv = 1.0f / lsum;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)(t0 + qi) * D + h * AT_DH);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      tc_fence_before();
      __syncthreads();  // TMEM (S, O) and the Q tile are reused by the next tile / head
    }
    __syncthreads();  // K / V^T are rewritten for the next head
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(ncols) : "memory");
  }
}

}  // namespace

extern "C" int acb_attention_varlen_tc(const void* qkv, const int* cu_seqlens, int B, int n_heads, int dh, int max_seqlen, float drop_p,
                                       long long seed, void* out, void* stream) {
  ACB_CHECK(qkv && cu_seqlens && out && B > 0 && n_heads > 0, "acb_attention_varlen_tc: bad arguments");
  ACB_CHECK(dh == AT_DH, "acb_attention_varlen_tc: head dim %d unsupported (16 only)", dh);
  ACB_CHECK(max_seqlen > 0 && max_seqlen <= 1024, "acb_attention_varlen_tc: max_seqlen %d exceeds the shared-memory K/V budget (1024 keys)", max_seqlen);
  ACB_CHECK(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0) && (n_heads * dh) % 8 == 0, "acb_attention_varlen_tc: alignment");
  ACB_CHECK(drop_p >= 0.0f && drop_p < 1.0f, "acb_attention_varlen_tc: bad dropout");
  const int ncap = ((max_seqlen + 15) / 16) * 16;
  const size_t smem = (size_t)2 * 128 * 16 + (size_t)2 * ncap * 16 + (size_t)(ncap / 8) * 256 + (size_t)2 * 8 * 128 * 16;
  // set on every call: the attribute is per device and the call is cheap (a process that touches a second GPU would otherwise
  // launch an unconfigured kernel there)
  ACB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<<<dim3(B, n_heads), AT_THREADS, smem, (cudaStream_t)stream>>>((const bf16*)qkv, cu_seqlens, n_heads, ncap, drop_p, acb_seed(seed),
                                                                     (bf16*)out, nullptr, nullptr);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

// the per-(sequence, head) kernel over a device-side list of sequences (attention_packed.cu: sequences longer than one tile)
int acb_attention_tc_long(const void* qkv, const int* cu_seqlens, const int* long_list, const int* n_long_dev, int grid_x, int n_heads,
                          int max_seqlen, float drop_p, long long seed, void* out, cudaStream_t st) {
  ACB_CHECK(max_seqlen > 0 && max_seqlen <= 1024, "acb_attention_packed: max_seqlen %d exceeds the shared-memory K/V budget (1024 keys)", max_seqlen);
  const int ncap = ((max_seqlen + 15) / 16) * 16;
  const size_t smem = (size_t)2 * 128 * 16 + (size_t)2 * ncap * 16 + (size_t)(ncap / 8) * 256 + (size_t)2 * 8 * 128 * 16;
  ACB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<<<dim3(grid_x, n_heads), AT_THREADS, smem, st>>>((const bf16*)qkv, cu_seqlens, n_heads, ncap, drop_p, acb_seed(seed), (bf16*)out,
                                                                    long_list, n_long_dev);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}
