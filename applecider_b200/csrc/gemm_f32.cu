// fp32 CUDA-core GEMM / implicit-GEMM conv1d with fused epilogue (parity path + tiny-N heads).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, fp32 FFMA accumulation.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct GemmF32Args {
  const float* A;
  const float* Bw;
  float* C;
  int M, N, K, lda, ldb, ldc;
  int conv_L, conv_Cin, conv_pad;
  const float* bias;
  int act;
  const float* res;
  int ldr;
  const float* gamma;
  int res_mode;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmF32Args p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  // per-thread load coordinates: 4 elements of A and 4 of B per k-tile
  int lm[4], lk[4];
  long long a_row_base[4];  // plain: m*lda ; conv: (b*L)*Cin
  int a_l[4];               // conv: l
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + i * 256;
    lk[i] = e & (BK - 1);
    lm[i] = e >> 4;
    const int m = m0 + lm[i];
    a_ok[i] = m < p.M;
    if (p.conv_L > 0) {
      const int b = m / p.conv_L;
      a_l[i] = m - b * p.conv_L;
      a_row_base[i] = (long long)b * p.conv_L * p.conv_Cin;
    } else {
      a_l[i] = 0;
      a_row_base[i] = (long long)m * p.lda;
    }
  }

  for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k0 + lk[i];
      float av = 0.0f;
      if (a_ok[i] && kk < p.K) {
        if (p.conv_L > 0) {
          const int tap = kk / p.conv_Cin;
          const int ci = kk - tap * p.conv_Cin;
          const int l = a_l[i] + tap - p.conv_pad;
          if (l >= 0 && l < p.conv_L) av = __ldg(p.A + a_row_base[i] + (long long)l * p.conv_Cin + ci);
        } else {
          av = __ldg(p.A + a_row_base[i] + kk);
        }
      }
      As[lk[i]][lm[i]] = av;
      const int n = n0 + lm[i];
      float bv = 0.0f;
      if (n < p.N && kk < p.K) bv = __ldg(p.Bw + (long long)n * p.ldb + kk);
      Bs[lk[i]][lm[i]] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      v = apply_act(v, p.act);
      if (p.res_mode == ACB_RES_ADD) {
        const float g = p.gamma ? __ldg(p.gamma + n) : 1.0f;
        v = __ldg(p.res + (long long)m * p.ldr + n) + g * v;
      } else if (p.res_mode == ACB_RES_MUL) {
        v = __ldg(p.res + (long long)m * p.ldr + n) * v;
      }
      p.C[(long long)m * p.ldc + n] = v;
    }
  }
}

}  // namespace

extern "C" int acb_gemm_f32(const float* A, const float* Bw, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                            int conv_L, int conv_Cin, int conv_pad, const float* bias, int act, const float* res,
                            int ldr, const float* gamma, int res_mode, void* stream) {
  ACB_CHECK(A && Bw && C, "acb_gemm_f32: null operand");
  ACB_CHECK(M >= 0 && N > 0 && K > 0, "acb_gemm_f32: bad shape M=%d N=%d K=%d", M, N, K);
  ACB_CHECK(res_mode == ACB_RES_NONE || res != nullptr, "acb_gemm_f32: res_mode set without res");
  if (conv_L > 0) {
    ACB_CHECK(conv_Cin > 0 && K % conv_Cin == 0 && M % conv_L == 0, "acb_gemm_f32: bad conv geometry");
  }
  if (M == 0) return ACB_OK;
  GemmF32Args p{A, Bw, C, M, N, K, lda, ldb, ldc, conv_L, conv_Cin, conv_pad, bias, act, res, ldr, gamma, res_mode};
  dim3 grid(cdiv(M, BM), cdiv(N, BN));
  ACB_CHECK(grid.y <= 65535, "acb_gemm_f32: N too large");
  gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}
