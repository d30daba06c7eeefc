// Fused ConvNeXt MLP block on tcgen05 (inference, bf16):   out = x + gamma * ( fc2( gelu( fc1(y) + b1 ) ) + b2 )
// timm ConvNeXtBlock tail (models/astrominn.py:12-17 -> timm convnext_tiny): y = LayerNorm(dwconv7(x)) [M, C], hidden 4C.
//
// Unfused this is two GEMMs with the [M, 4C] hidden activation written to and read back from HBM (stage 0 of a 4096-cutout
// batch: 921,600 x 384 bf16 = 708 MB each way, per block).  Here one CTA owns 128 rows and walks the hidden dimension in
// chunks of HC columns:
//     acc1[128 x HC]  = Y_tile W1[chunk]^T            (tcgen05, accumulator in TMEM, double buffered)
//     H_chunk         = bf16( gelu(acc1 + b1) )       (epilogue warps: TMEM -> registers -> SWIZZLE_128B smem tile)
//     acc2[128 x C]  += H_chunk W2[:, chunk]^T        (tcgen05, A operand = the smem tile just written)
// so the hidden activation only ever exists as one chunk in shared memory.  The final epilogue adds b2, scales by gamma,
// adds the residual and stores bf16.
//   warp 0   TMA producer: Y tile once, then W1 / W2 chunks (L2-resident weights) through double-buffered rings
//   warp 1   MMA issuer: GEMM1 of chunk j is issued BEFORE GEMM2 of chunk j-1, so the tensor pipe works while the epilogue
//            warps are still busy with chunk j-1
//   warps 2+ epilogue: four warps per 32 hidden columns (one per TMEM lane quarter); thread = row
#include "tc_common.cuh"

using namespace tc;

namespace {

struct MlpArgs {
  long long M;
  const int* rows_dev;  // optional device-side row count (capacity-sized token matrices): tiles past it exit at once
  const float* b1;
  const float* b2;
  const float* gamma;
  const bf16* res;
  bf16* out;
};

template <int C, int HC>
struct MlpCfg {
  static constexpr int KB1 = (C + 63) / 64;          // 64-wide K blocks of GEMM1 (K = C; the last one may be half empty)
  static constexpr int KS1 = C / 16;                 // UMMA K steps of GEMM1
  static constexpr int NJ = 4 * C / HC;              // hidden chunks
  static constexpr int KB2 = HC / 64;                // K blocks of GEMM2 per chunk
  static constexpr int NEW = HC / 32;                // epilogue warps per TMEM lane quarter
  static constexpr int NE_WARPS = 4 * NEW;
  static constexpr int THREADS = 64 + 32 * NE_WARPS;
  static constexpr int NW2 = (C == 192) ? 2 : 1;     // W2 ring depth (what fits next to the other tiles)
  static constexpr uint32_t Y_BYTES = KB1 * 16384;
  static constexpr uint32_t W1_BYTES = KB1 * HC * 128;
  static constexpr uint32_t W2_BYTES = KB2 * C * 128;
  static constexpr uint32_t H_BYTES = KB2 * 16384;
  static constexpr uint32_t SMEM = Y_BYTES + 2 * W1_BYTES + NW2 * W2_BYTES + 2 * H_BYTES + 1024 + (4 * C + 2 * C) * 4;
  static_assert(C % 32 == 0 && HC % 64 == 0 && (4 * C) % HC == 0, "shape");
  static_assert(2 * HC + C <= 512, "TMEM");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

__device__ __forceinline__ void tmem_ld16b(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int C, int HC, int ACT>
__global__ void __launch_bounds__(MlpCfg<C, HC>::THREADS, 1) mlp_block_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                             const __grid_constant__ CUtensorMap tmW1,
                                                                             const __grid_constant__ CUtensorMap tmW2,
                                                                             const __grid_constant__ MlpArgs p) {
  using G = MlpCfg<C, HC>;
  constexpr uint32_t IDESC1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  // barriers: 0 y_full | 1,2 w1_full | 3,4 w1_empty | 5,6 w2_full | 7,8 w2_empty | 9,10 a1_full | 11,12 a1_free | 13,14 h_ready |
  //           15,16 h_free | 17 a2_full
  __shared__ __align__(8) uint64_t bars[18];
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * 128;
  if (p.rows_dev && m0 >= (long long)(*p.rows_dev)) return;  // uniform, before any barrier / TMEM allocation
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t aY = base, aW1 = aY + G::Y_BYTES, aW2 = aW1 + 2 * G::W1_BYTES, aH = aW2 + G::NW2 * G::W2_BYTES;
  uint8_t* sH = gen + (aH - base);
  float* s_b1 = reinterpret_cast<float*>(gen + (aH - base) + 2 * G::H_BYTES);
  float* s_b2 = s_b1 + 4 * C;
  float* s_g = s_b2 + C;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto B = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 18; ++i) mbar_init(B(i), (i >= 11 && i <= 14) ? (uint32_t)G::NE_WARPS : 1u);  // a1_free, h_ready: one arrival per epilogue warp
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4 * C; i += G::THREADS) s_b1[i] = __ldg(p.b1 + i);
  for (int i = threadIdx.x; i < C; i += G::THREADS) { s_b2[i] = __ldg(p.b2 + i); s_g[i] = __ldg(p.gamma + i); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_a1 = tmem_holder;               // [2][HC] columns
  const uint32_t tmem_a2 = tmem_holder + 2u * HC;     // [C] columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      mbar_expect_tx(B(0), G::Y_BYTES);
      for (int kb = 0; kb < G::KB1; ++kb) tma_load_2d(aY + kb * 16384, &tmY, kb * 64, (int)m0, B(0));
    }
    __syncwarp();
    for (int j = 0; j < G::NJ; ++j) {
      const int b = j & 1, b2 = j % G::NW2;
      mbar_wait(B(3 + b), (uint32_t)(((j >> 1) & 1) ^ 1));
      if (elect_one_sync()) {
        mbar_expect_tx(B(1 + b), G::W1_BYTES);
        for (int kb = 0; kb < G::KB1; ++kb) tma_load_2d(aW1 + b * G::W1_BYTES + kb * (HC * 128), &tmW1, kb * 64, j * HC, B(1 + b));
      }
      __syncwarp();
      mbar_wait(B(7 + b2), (uint32_t)((((j / G::NW2)) & 1) ^ 1));
      if (elect_one_sync()) {
        mbar_expect_tx(B(5 + b2), G::W2_BYTES);
        for (int kb = 0; kb < G::KB2; ++kb) tma_load_2d(aW2 + b2 * G::W2_BYTES + kb * (C * 128), &tmW2, j * HC + kb * 64, 0, B(5 + b2));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(B(0), 0);
    for (int j = 0; j <= G::NJ; ++j) {
      if (j < G::NJ) {
        const int b = j & 1;
        mbar_wait(B(1 + b), (uint32_t)((j >> 1) & 1));
        mbar_wait(B(11 + b), (uint32_t)(((j >> 1) & 1) ^ 1));
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < G::KS1; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            umma_bf16(tmem_a1 + (uint32_t)(b * HC), make_smem_desc(aY + kb * 16384) + 2 * kk,
                      make_smem_desc(aW1 + b * G::W1_BYTES + kb * (HC * 128)) + 2 * kk, IDESC1, ks > 0 ? 1u : 0u);
          }
          umma_commit(B(3 + b));   // W1 buffer free
          umma_commit(B(9 + b));   // acc1 ready
        }
        __syncwarp();
      }
      if (j > 0) {
        const int i = j - 1, b = i & 1, b2 = i % G::NW2;
        mbar_wait(B(13 + b), (uint32_t)((i >> 1) & 1));
        mbar_wait(B(5 + b2), (uint32_t)((i / G::NW2) & 1));
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < HC / 16; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            umma_bf16(tmem_a2, make_smem_desc(aH + b * G::H_BYTES + kb * 16384) + 2 * kk,
                      make_smem_desc(aW2 + b2 * G::W2_BYTES + kb * (C * 128)) + 2 * kk, IDESC2, (i > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(B(15 + b));   // H buffer free
          umma_commit(B(7 + b2));   // W2 buffer free
        }
        __syncwarp();
      }
    }
    if (elect_one_sync()) umma_commit(B(17));
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may touch
    const int cg = e >> 2;           // 32-column group of the hidden chunk
    const int r = q * 32 + lane;     // tile row
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int j = 0; j < G::NJ; ++j) {
      const int b = j & 1;
      mbar_wait(B(9 + b), (uint32_t)((j >> 1) & 1));
      mbar_wait(B(15 + b), (uint32_t)(((j >> 1) & 1) ^ 1));
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32(tmem_a1 + (uint32_t)(b * HC) + lane_addr + (uint32_t)(cg * 32), raw);
      const float* bb = s_b1 + j * HC + cg * 32;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 b2v = *reinterpret_cast<const float2*>(bb + i);
        const float2 uu = __fadd2_rn(make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), b2v);
        const float2 aa = ACT == ACB_ACT_RELU ? make_float2(fmaxf(uu.x, 0.0f), fmaxf(uu.y, 0.0f)) : gelu_bf16x2(uu);
        __nv_bfloat162 hh = __floats2bfloat162_rn(aa.x, aa.y);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
      }
      // H tile: K block (cg*32)/64, 16-byte chunks c0..c0+3 of row r, SWIZZLE_128B
      uint8_t* rowp = sH + b * G::H_BYTES + ((cg * 32) >> 6) * 16384 + r * 128;
      const int c0 = ((cg * 32) & 63) >> 3;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        *reinterpret_cast<uint4*>(rowp + (((c0 + t) ^ (r & 7)) << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(B(11 + b));  // acc1 buffer may be overwritten
        mbar_arrive(B(13 + b));  // H chunk is complete
      }
    }
    // ---- final epilogue: out = res + gamma * (acc2 + b2) ----
    mbar_wait_sleep(B(17), 0);
    tc_fence_after();
    const long long m = m0 + r;
    for (int g = cg; g < C / 16; g += G::NEW) {
      uint32_t o[16];
      tmem_ld16b(tmem_a2 + lane_addr + (uint32_t)(g * 16), o);
      if (m < p.M) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + m * C + g * 16);
        const uint4 r0 = rp[0], r1 = rp[1];
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = g * 16 + 2 * i;
          const float v0 = fmaf(s_g[c], __uint_as_float(o[2 * i]) + s_b2[c], __uint_as_float(rw[i] << 16));
          const float v1 = fmaf(s_g[c + 1], __uint_as_float(o[2 * i + 1]) + s_b2[c + 1], __uint_as_float(rw[i] & 0xffff0000u));
          __nv_bfloat162 hh = __floats2bfloat162_rn(v0, v1);
          w[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        uint4* op = reinterpret_cast<uint4*>(p.out + m * C + g * 16);
        op[0] = make_uint4(w[0], w[1], w[2], w[3]);
        op[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_holder), "r"(512u) : "memory");
  }
}

template <int C, int HC, int ACT>
int launch_mlp(const void* y, const void* w1, const void* w2, const MlpArgs& args, cudaStream_t st) {
  using G = MlpCfg<C, HC>;
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
  ACB_CHECK(enc != nullptr, "acb_convnext_mlp_bf16: cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmY, tmW1, tmW2;
  auto mk = [&](CUtensorMap* tm, const void* ptr, long long rows, int cols, int box_rows) -> CUresult {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  ACB_CHECK(mk(&tmY, y, args.M, C, 128) == CUDA_SUCCESS, "acb_convnext_mlp_bf16: tensor map (y) failed");
  ACB_CHECK(mk(&tmW1, w1, 4 * C, C, HC) == CUDA_SUCCESS, "acb_convnext_mlp_bf16: tensor map (fc1 weight) failed");
  ACB_CHECK(mk(&tmW2, w2, C, 4 * C, C) == CUDA_SUCCESS, "acb_convnext_mlp_bf16: tensor map (fc2 weight) failed");
  auto k = mlp_block_kernel<C, HC, ACT>;
  ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
  k<<<(unsigned)((args.M + 127) / 128), G::THREADS, G::SMEM, st>>>(tmY, tmW1, tmW2, args);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // namespace

extern "C" int acb_convnext_mlp_bf16(const void* y, const void* res, const void* w1, const float* b1, const void* w2, const float* b2,
                                     const float* gamma, void* out, long long M, int C, void* stream) {
  ACB_CHECK(y && res && w1 && b1 && w2 && b2 && gamma && out && M > 0, "acb_convnext_mlp_bf16: bad arguments");
  ACB_CHECK((((uintptr_t)y | (uintptr_t)res | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)out) & 15) == 0, "acb_convnext_mlp_bf16: 16-byte alignment");
  MlpArgs args{M, nullptr, b1, b2, gamma, (const bf16*)res, (bf16*)out};
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 96) return launch_mlp<96, 128, ACB_ACT_GELU>(y, w1, w2, args, st);
  if (C == 192) return launch_mlp<192, 64, ACB_ACT_GELU>(y, w1, w2, args, st);
  acb_set_error("acb_convnext_mlp_bf16: C = %d is not fused (96 and 192 are; wider stages hold too few rows to matter)", C);
  return ACB_ERR_UNSUPPORTED;
}

// Transformer feed-forward block (nn.TransformerEncoderLayer, HyraxBaselineCLS.py:26-33: linear1 128 -> 512, ReLU, linear2 512 -> 128,
// residual add) on the same kernel: out = x + linear2(relu(linear1(x) + b1)) + b2; `ones` = a vector of C ones (the layer-scale slot).
extern "C" int acb_ffn_relu_bf16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* ones, void* out,
                                 long long M, int C, const int* rows_dev, void* stream) {
  ACB_CHECK(x && w1 && b1 && w2 && b2 && ones && out && M > 0, "acb_ffn_relu_bf16: bad arguments");
  ACB_CHECK((((uintptr_t)x | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)out) & 15) == 0, "acb_ffn_relu_bf16: 16-byte alignment");
  ACB_CHECK(C == 128, "acb_ffn_relu_bf16: d_model = %d is not supported (128 with a 4x feed-forward width is)", C);
  MlpArgs args{M, rows_dev, b1, b2, ones, (const bf16*)x, (bf16*)out};
  return launch_mlp<128, 128, ACB_ACT_RELU>(x, w1, w2, args, (cudaStream_t)stream);
}
