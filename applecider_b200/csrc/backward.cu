// Backward / training kernels: generic strided fp32 GEMM with split-K (dgrad / wgrad / conv wgrad),
// column reductions, activation / LayerNorm / attention / depthwise-conv / pooling backward, embedding and
// MoE backward, losses, dropout.  Elementwise kernels are dtype-tagged (f32 | bf16 IO, fp32 math).
#include <algorithm>

#include "common.cuh"

namespace {

inline unsigned grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ================================ generic strided GEMM (fp32 accumulate) ==============================
// C[m,n] (+)= sum_k A(m,k) * B(n,k);  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk] (dtype-tagged).
// convT != 0: A is the transposed im2col of a channels-last signal X[nb, L, Cin]:
//   m = tap*Cin + ci, k = b*L + l  ->  X[b, l + tap - pad, ci] (zero outside [0, L)).
// grid.z splits K; with splits > 1 (or accumulate) results are atomically added into fp32 C.
struct GemmExArgs {
  const void* A; const void* B; float* C;
  int a_dt, b_dt;
  int M, N, K;
  long long sam, sak, sbn, sbk;
  int ldc;
  int convT, conv_L, conv_Cin, conv_pad;
  int k_chunk, atomic;
};

__global__ void __launch_bounds__(256) gemm_ex_kernel(const GemmExArgs p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * p.k_chunk;
  const int k_end = min(p.K, k_begin + p.k_chunk);
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      // A tile: when sak == 1 walk k fastest, else walk m fastest (coalescing)
      int lk, lm;
      if (p.sak == 1 && !p.convT) { lk = e & (BK - 1); lm = e >> 4; } else { lm = e & (BM - 1); lk = e >> 6; }
      const int m = m0 + lm, kk = k0 + lk;
      float av = 0.0f;
      if (m < p.M && kk < k_end) {
        if (p.convT) {
          const int tap = m / p.conv_Cin, ci = m - tap * p.conv_Cin;
          const int b = kk / p.conv_L, l = kk - b * p.conv_L + tap - p.conv_pad;
          if (l >= 0 && l < p.conv_L) av = ld_any(p.A, ((long long)b * p.conv_L + l) * p.conv_Cin + ci, p.a_dt);
        } else {
          av = ld_any(p.A, (long long)m * p.sam + (long long)kk * p.sak, p.a_dt);
        }
      }
      As[lk][lm] = av;
      int bk, bn_;
      if (p.sbk == 1) { bk = e & (BK - 1); bn_ = e >> 4; } else { bn_ = e & (BN - 1); bk = e >> 6; }
      const int n = n0 + bn_, kb = k0 + bk;
      float bv = 0.0f;
      if (n < p.N && kb < k_end) bv = ld_any(p.B, (long long)n * p.sbn + (long long)kb * p.sbk, p.b_dt);
      Bs[bk][bn_] = bv;
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
      float* c = p.C + (long long)m * p.ldc + n;
      if (p.atomic) atomicAdd(c, acc[i][j]); else *c = acc[i][j];
    }
  }
}

// ================================ reductions / elementwise ==========================================
// out[n] (+)= sum_m a[m,n] * (b ? b[m,n] : 1)
__global__ void __launch_bounds__(256) colsum_kernel(const void* a, int a_dt, const void* b, int b_dt, long long M, int N,
                                                     long long ld, int rows_per_block, float* out) {
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;  // 8 row lanes
  float s = 0.0f;
  if (n < N) {
    for (long long m = r0 + rl; m < r1; m += 8) {
      float v = ld_any(a, m * ld + n, a_dt);
      if (b) v *= ld_any(b, m * ld + n, b_dt);
      s += v;
    }
  }
  __shared__ float sh[8][33];
  sh[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && n < N) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x & 31];
    atomicAdd(out + n, t);
  }
}

// bf16 operands, N % 8 == 0, ld % 8 == 0, N <= 2048: thread = (row lane, group of 8 columns), whole rows move as consecutive 16-byte
// loads; the row lanes meet in a shared-memory accumulator, one global atomic per column and CTA
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, long long M, int N8,
                                                            long long ld8, int rows_per_block, float* __restrict__ out) {
  __shared__ float acc[2048];
  for (int i = threadIdx.x; i < N8 * 8; i += 256) acc[i] = 0.0f;
  __syncthreads();
  const int lanes = 256 / N8;  // row lanes (>= 1); the last 256 % N8 threads idle
  const int rl = threadIdx.x / N8, cg = threadIdx.x - rl * N8;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
  if (rl < lanes) {
    float s[8] = {};
    for (long long m = r0 + rl; m < r1; m += lanes) {
      const uint4 v = __ldcs(a + m * ld8 + cg);
      const uint32_t* vw = reinterpret_cast<const uint32_t*>(&v);
      if (b) {
        const uint4 u = __ldcs(b + m * ld8 + cg);
        const uint32_t* uw = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vw[j]));
          const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&uw[j]));
          s[2 * j] = fmaf(f.x, g.x, s[2 * j]);
          s[2 * j + 1] = fmaf(f.y, g.y, s[2 * j + 1]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vw[j]));
          s[2 * j] += f.x;
          s[2 * j + 1] += f.y;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&acc[cg * 8 + j], s[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N8 * 8; i += 256) atomicAdd(out + i, acc[i]);
}

__global__ void act_fwd_kernel(const void* x, int x_dt, void* y, int y_dt, int act, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st_any(y, i, y_dt, apply_act(ld_any(x, i, x_dt), act));
}

// bf16 everywhere, n % 8 == 0, ReLU / GELU (the two activations of the big hidden layers): 16 bytes per thread and step
__global__ void __launch_bounds__(256) act_bwd_bf16x8_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ dx, int act,
                                                             long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 gv = __ldcs(dy + i);
    const uint4 xv = __ldcs(x + i);
    uint32_t* gw = reinterpret_cast<uint32_t*>(&gv);
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[j]));
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[j]));
      const float d0 = act == ACB_ACT_RELU ? (v.x > 0.0f ? 1.0f : 0.0f) : gelu_bf16_grad(v.x);
      const float d1 = act == ACB_ACT_RELU ? (v.y > 0.0f ? 1.0f : 0.0f) : gelu_bf16_grad(v.y);
      const __nv_bfloat162 o = __floats2bfloat162_rn(g.x * d0, g.y * d1);
      gw[j] = *reinterpret_cast<const uint32_t*>(&o);
    }
    dx[i] = gv;
  }
}

// dx = dy * act'(x)   (x = pre-activation)
__global__ void act_bwd_kernel(const void* dy, int dy_dt, const void* x, int x_dt, void* dx, int dx_dt, int act, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = ld_any(dy, i, dy_dt), v = ld_any(x, i, x_dt);
    float d;
    switch (act) {
      case ACB_ACT_RELU: d = v > 0.0f ? 1.0f : 0.0f; break;
      case ACB_ACT_GELU: d = (x_dt == ACB_BF16 && dx_dt == ACB_BF16) ? gelu_bf16_grad(v) : gelu_erf_grad(v); break;  // bf16 forward = tanh form
      case ACB_ACT_TANH: { const float t = tanhf(v); d = 1.0f - t * t; break; }
      case ACB_ACT_SIGMOID: { const float s = sigmoidf_(v); d = s * (1.0f - s); break; }
      case ACB_ACT_SOFTPLUS: d = v > 20.0f ? 1.0f : sigmoidf_(v); break;
      default: d = 1.0f;
    }
    st_any(dx, i, dx_dt, g * d);
  }
}

// op 0: y = a + b; 1: y = a * b; 2: y = a + g[col]*b; 3: y = g[col]*a; 4: y = a*s0 + b*s1 (scalars); 5: y = a * g[0]
__global__ void ew_kernel(const void* a, int a_dt, const void* b, int b_dt, const float* g, void* y, int y_dt, int op, int C,
                          float s0, float s1, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float av = ld_any(a, i, a_dt);
    const float bv = b ? ld_any(b, i, b_dt) : 0.0f;
    float r;
    switch (op) {
      case 0: r = av + bv; break;
      case 1: r = av * bv; break;
      case 2: r = av + g[i % C] * bv; break;
      case 3: r = g[i % C] * av; break;
      case 5: r = av * g[0]; break;
      default: r = av * s0 + bv * s1;
    }
    st_any(y, i, y_dt, r);
  }
}

// a, b, y bf16, n % 8 == 0 (ops 2 / 3: C % 8 == 0 as well): 16 bytes per thread and step
__global__ void __launch_bounds__(256) ew_bf16x8_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const float* __restrict__ g,
                                                        uint4* __restrict__ y, int op, int C, float s0, float s1, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 av = a[i];
    const uint4 bv = b ? b[i] : make_uint4(0u, 0u, 0u, 0u);
    uint32_t* aw = reinterpret_cast<uint32_t*>(&av);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(&bv);
    float gv[8];
    if (op == 2 || op == 3) {
      const float4* gp = reinterpret_cast<const float4*>(g + (int)((i * 8) % C));
      const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
      gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[j]));
      const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[j]));
      float2 r;
      if (op == 0) r = make_float2(fa.x + fb.x, fa.y + fb.y);
      else if (op == 1) r = make_float2(fa.x * fb.x, fa.y * fb.y);
      else if (op == 2) r = make_float2(fa.x + gv[2 * j] * fb.x, fa.y + gv[2 * j + 1] * fb.y);
      else if (op == 3) r = make_float2(gv[2 * j] * fa.x, gv[2 * j + 1] * fa.y);
      else r = make_float2(fa.x * s0 + fb.x * s1, fa.y * s0 + fb.y * s1);
      const __nv_bfloat162 o = __floats2bfloat162_rn(r.x, r.y);
      aw[j] = *reinterpret_cast<const uint32_t*>(&o);
    }
    y[i] = av;
  }
}

__global__ void copy2d_kernel(const void* src, int s_dt, long long lds, void* dst, int d_dt, long long ldd, long long rows, int cols) {
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    st_any(dst, r * ldd + c, d_dt, ld_any(src, r * lds + c, s_dt));
  }
}

__global__ void gather_cols_kernel(const float* X, int ldx, const int* cols, int n, float* Y, long long rows) {
  const long long tot = rows * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    Y[i] = X[r * ldx + cols[i - r * n]];
  }
}

// dW[co,ci,tap] = G[(tap*Cin + ci)*Cout + co]   (G = output of the convT wgrad GEMM)
__global__ void unpack_conv_wgrad_kernel(const float* G, float* dW, int Cout, int Cin, int k) {
  const long long tot = (long long)Cout * Cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % k);
    const long long t = i / k;
    const int ci = (int)(t % Cin);
    const int co = (int)(t / Cin);
    dW[i] = G[((long long)tap * Cin + ci) * Cout + co];
  }
}
// dgrad weights: out[ci*(k*Cout) + tap*Cout + co] = w[co, ci, k-1-tap]
__global__ void pack_conv_dgrad_kernel(const float* w, void* out, int odt, int Cout, int Cin, int k) {
  const long long tot = (long long)Cout * Cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long t = i / Cout;
    const int tap = (int)(t % k);
    const int ci = (int)(t / k);
    st_any(out, i, odt, w[((long long)co * Cin + ci) * k + (k - 1 - tap)]);
  }
}

// ================================ LayerNorm backward ================================================
// warp per row; dx = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*w; dw += dy*xhat, db += dy (atomics per block)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const void* x, int x_dt, const void* dy, int dy_dt, const float* w,
                                                            const float* bias, int gelu, void* dx, int dx_dt, float* dw, float* db,
                                                            long long rows, int C, float eps, int rows_per_warp) {
  extern __shared__ float shacc[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) shacc[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long row0 = ((long long)blockIdx.x * nw + wid) * rows_per_warp;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const long long row = row0 + rr;
    if (row >= rows) break;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += ld_any(x, row * C + c, x_dt);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float d = ld_any(x, row * C + c, x_dt) - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    float sg = 0.0f, sgx = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (ld_any(x, row * C + c, x_dt) - mean) * rstd;
      float dyv = ld_any(dy, row * C + c, dy_dt);
      if (gelu) dyv *= gelu_erf_grad(fmaf(xh, w[c], bias[c]));  // y = gelu(LN(x)): chain through the activation
      const float g = dyv * w[c];
      sg += g;
      sgx += g * xh;
    }
    sg = warp_sum(sg) / (float)C;
    sgx = warp_sum(sgx) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (ld_any(x, row * C + c, x_dt) - mean) * rstd;
      float dyv = ld_any(dy, row * C + c, dy_dt);
      if (gelu) dyv *= gelu_erf_grad(fmaf(xh, w[c], bias[c]));
      st_any(dx, row * C + c, dx_dt, rstd * (dyv * w[c] - sg - xh * sgx));
      atomicAdd(&shacc[c], dyv * xh);
      atomicAdd(&shacc[C + c], dyv);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dw + i, shacc[i]);
    atomicAdd(db + i, shacc[C + i]);
  }
}

// Register-resident variant for C = 32*VEC*NCH <= 768 and one dtype for x/dy/dx: every row is read ONCE (vector
// loads), dw/db partial sums stay in registers over the warp's rows, then one shared-memory pass + one global
// atomic per channel and CTA.
template <typename T, int NCH, int VEC>
__global__ void __launch_bounds__(256) layernorm_bwd_reg_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ w,
                                                                const float* __restrict__ bias, int gelu, T* __restrict__ dx,
                                                                float* __restrict__ dw, float* __restrict__ db, long long rows, float eps,
                                                                int rows_per_warp) {
  constexpr int C = 32 * VEC * NCH, NE = NCH * VEC;
  __shared__ float shacc[2 * C];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) shacc[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float wv[NE], bv[NE], aw[NE], ab[NE];
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      wv[k * VEC + e] = w[(k * 32 + lane) * VEC + e];
      bv[k * VEC + e] = gelu ? bias[(k * 32 + lane) * VEC + e] : 0.0f;
      aw[k * VEC + e] = 0.0f;
      ab[k * VEC + e] = 0.0f;
    }
  const long long row0 = ((long long)blockIdx.x * nw + wid) * rows_per_warp;
  constexpr bool PREFETCH = VEC == 2 && sizeof(T) == 2;  // software prefetch of the next row: twice the bytes in flight per warp
  uint32_t px[PREFETCH ? NCH : 1], pg[PREFETCH ? NCH : 1];
  if constexpr (PREFETCH) {
    if (row0 < rows) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        px[k] = *reinterpret_cast<const uint32_t*>(x + row0 * C + (k * 32 + lane) * 2);
        pg[k] = *reinterpret_cast<const uint32_t*>(dy + row0 * C + (k * 32 + lane) * 2);
      }
    }
  }
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const long long row = row0 + rr;
    if (row >= rows) break;
    float xv[NE], gv[NE];
    const T* xr = x + row * C;
    const T* gr = dy + row * C;
    if constexpr (PREFETCH) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&px[k]));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pg[k]));
        xv[k * 2] = a.x; xv[k * 2 + 1] = a.y; gv[k * 2] = b.x; gv[k * 2 + 1] = b.y;
      }
      if (rr + 1 < rows_per_warp && row + 1 < rows) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          px[k] = *reinterpret_cast<const uint32_t*>(xr + C + (k * 32 + lane) * 2);
          pg[k] = *reinterpret_cast<const uint32_t*>(gr + C + (k * 32 + lane) * 2);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
      if constexpr (PREFETCH) {
        (void)c0;
      } else if constexpr (VEC == 2 && sizeof(T) == 2) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xr + c0));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gr + c0));
        xv[k * 2] = a.x; xv[k * 2 + 1] = a.y; gv[k * 2] = b.x; gv[k * 2 + 1] = b.y;
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          xv[k * VEC + e] = to_f<T>(xr[c0 + e]);
          gv[k * VEC + e] = to_f<T>(gr[c0 + e]);
        }
      }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) s += xv[i];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) { xv[i] -= mean; q += xv[i] * xv[i]; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    float sg = 0.0f, sgx = 0.0f;
    if constexpr (VEC == 2 && sizeof(T) == 2) {
      // bf16 rows: the per-element arithmetic on the packed fp32 pipe (two channels per instruction; the kernel is issue-bound: ncu 81 %)
      float2 sg2 = make_float2(0.0f, 0.0f), sgx2 = make_float2(0.0f, 0.0f);
      const float2 rs2 = make_float2(rstd, rstd);
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const float2 w2 = make_float2(wv[2 * k], wv[2 * k + 1]);
        const float2 xh = __fmul2_rn(make_float2(xv[2 * k], xv[2 * k + 1]), rs2);  // xhat
        float2 g2 = make_float2(gv[2 * k], gv[2 * k + 1]);
        if (gelu) {
          const float2 pre = __ffma2_rn(xh, w2, make_float2(bv[2 * k], bv[2 * k + 1]));
          g2 = __fmul2_rn(g2, make_float2(gelu_bf16_grad(pre.x), gelu_bf16_grad(pre.y)));
        }
        const float2 gw = __fmul2_rn(g2, w2);
        sg2 = __fadd2_rn(sg2, gw);
        sgx2 = __ffma2_rn(gw, xh, sgx2);
        const float2 a2 = __ffma2_rn(g2, xh, make_float2(aw[2 * k], aw[2 * k + 1]));
        const float2 b2 = __fadd2_rn(make_float2(ab[2 * k], ab[2 * k + 1]), g2);
        aw[2 * k] = a2.x; aw[2 * k + 1] = a2.y;
        ab[2 * k] = b2.x; ab[2 * k + 1] = b2.y;
        xv[2 * k] = xh.x; xv[2 * k + 1] = xh.y;
        gv[2 * k] = gw.x; gv[2 * k + 1] = gw.y;  // from here on gv holds g * w
      }
      sg = sg2.x + sg2.y;
      sgx = sgx2.x + sgx2.y;
    } else {
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        xv[i] *= rstd;  // xhat
        if (gelu) gv[i] *= (sizeof(T) == 2 ? gelu_bf16_grad(fmaf(xv[i], wv[i], bv[i])) : gelu_erf_grad(fmaf(xv[i], wv[i], bv[i])));
        const float g = gv[i] * wv[i];
        sg += g;
        sgx += g * xv[i];
        aw[i] = fmaf(gv[i], xv[i], aw[i]);
        ab[i] += gv[i];
      }
    }
    sg = warp_sum(sg) / (float)C;
    sgx = warp_sum(sgx) / (float)C;
    T* dr = dx + row * C;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
      if constexpr (VEC == 2 && sizeof(T) == 2) {
        // rstd * (g w - sg - xhat sgx)
        const float2 t2 = __ffma2_rn(make_float2(-xv[k * 2], -xv[k * 2 + 1]), make_float2(sgx, sgx),
                                     __fadd2_rn(make_float2(gv[k * 2], gv[k * 2 + 1]), make_float2(-sg, -sg)));
        const float2 o2 = __fmul2_rn(make_float2(rstd, rstd), t2);
        *reinterpret_cast<__nv_bfloat162*>(dr + c0) = __floats2bfloat162_rn(o2.x, o2.y);
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int i = k * VEC + e;
          dr[c0 + e] = from_f<T>(rstd * (gv[i] * wv[i] - sg - xv[i] * sgx));
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int c = (k * 32 + lane) * VEC + e;
      atomicAdd(&shacc[c], aw[k * VEC + e]);
      atomicAdd(&shacc[C + c], ab[k * VEC + e]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dw + i, shacc[i]);
    atomicAdd(db + i, shacc[C + i]);
  }
}

// Wide rows (C = NPT x 512 >= 1024: the last SpectraNet blocks): a CTA walks rows_per_cta rows, thread t owns the column pairs
// (k 256 + t) of every row, so x / dy / dx move as coalesced 4-byte accesses, the d(weight) / d(bias) partial sums live in registers
// for the whole walk and the next row is in flight while the current one is reduced (three block reductions per row).
template <int NPT>
__global__ void __launch_bounds__(256) layernorm_bwd_wide_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                                 const float* __restrict__ w, const float* __restrict__ bias, int gelu,
                                                                 bf16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                                 long long rows, float eps, int rows_per_cta) {
  constexpr int C = NPT * 512, NE = NPT * 2;
  __shared__ float red[2][2][8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int phase = 0;
  auto block_sum2 = [&](float& a, float& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      red[phase][0][wid] = a;
      red[phase][1][wid] = b;
    }
    __syncthreads();
    a = 0.0f, b = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a += red[phase][0][i];
      b += red[phase][1][i];
    }
    phase ^= 1;  // the buffer written two reductions ago is free again: every thread has passed the barrier in between
  };
  float wv[NE], bv[NE], aw[NE], ab[NE];
#pragma unroll
  for (int k = 0; k < NPT; ++k)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = (k * 256 + tid) * 2 + e;
      wv[k * 2 + e] = w[c];
      bv[k * 2 + e] = gelu ? bias[c] : 0.0f;
      aw[k * 2 + e] = 0.0f;
      ab[k * 2 + e] = 0.0f;
    }
  const long long row0 = (long long)blockIdx.x * rows_per_cta, row1 = min(rows, row0 + rows_per_cta);
  uint32_t px[NPT], pg[NPT];
  if (row0 < row1) {
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      px[k] = __ldcs(reinterpret_cast<const uint32_t*>(x + row0 * C) + k * 256 + tid);
      pg[k] = __ldcs(reinterpret_cast<const uint32_t*>(dy + row0 * C) + k * 256 + tid);
    }
  }
  for (long long row = row0; row < row1; ++row) {
    float xv[NE], gv[NE];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&px[k]));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pg[k]));
      xv[k * 2] = a.x; xv[k * 2 + 1] = a.y; gv[k * 2] = b.x; gv[k * 2 + 1] = b.y;
    }
    if (row + 1 < row1) {
#pragma unroll
      for (int k = 0; k < NPT; ++k) {
        px[k] = __ldcs(reinterpret_cast<const uint32_t*>(x + (row + 1) * C) + k * 256 + tid);
        pg[k] = __ldcs(reinterpret_cast<const uint32_t*>(dy + (row + 1) * C) + k * 256 + tid);
      }
    }
    float s = 0.0f, dummy = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) s += xv[i];
    block_sum2(s, dummy);
    const float mean = s / (float)C;
    float q = 0.0f;
    dummy = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) { xv[i] -= mean; q += xv[i] * xv[i]; }
    block_sum2(q, dummy);
    const float rstd = rsqrtf(q / (float)C + eps);
    float sg = 0.0f, sgx = 0.0f;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      xv[i] *= rstd;  // xhat
      if (gelu) gv[i] *= gelu_bf16_grad(fmaf(xv[i], wv[i], bv[i]));
      const float g = gv[i] * wv[i];
      sg += g;
      sgx += g * xv[i];
      aw[i] = fmaf(gv[i], xv[i], aw[i]);
      ab[i] += gv[i];
    }
    block_sum2(sg, sgx);
    sg /= (float)C;
    sgx /= (float)C;
    uint32_t* dr = reinterpret_cast<uint32_t*>(dx + row * C);
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const float o0 = rstd * (gv[k * 2] * wv[k * 2] - sg - xv[k * 2] * sgx);
      const float o1 = rstd * (gv[k * 2 + 1] * wv[k * 2 + 1] - sg - xv[k * 2 + 1] * sgx);
      const __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
      dr[k * 256 + tid] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
#pragma unroll
  for (int k = 0; k < NPT; ++k)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = (k * 256 + tid) * 2 + e;
      atomicAdd(dw + c, aw[k * 2 + e]);
      atomicAdd(db + c, ab[k * 2 + e]);
    }
}

static bool launch_ln_bwd_wide(const void* x, const void* dy, const float* w, const float* bias, int gelu, void* dx, float* dw, float* db,
                               long long rows, int C, float eps, cudaStream_t st) {
  const int rpc = (int)std::max<long long>(1, std::min<long long>(64, rows / (148 * 8 * 2)));
  const unsigned grid = (unsigned)((rows + rpc - 1) / rpc);
#define LNW(NPT)                                                                                                                  \
  layernorm_bwd_wide_kernel<NPT><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)dy, w, bias, gelu, (bf16*)dx, dw, db, rows, eps, rpc); \
  return true
  switch (C) {
    case 1024: LNW(2);
    case 1536: LNW(3);
    case 2048: LNW(4);
    case 3072: LNW(6);
    default: return false;
  }
#undef LNW
}

template <typename T>
static bool launch_ln_bwd_reg(const void* x, const void* dy, const float* w, const float* bias, int gelu, void* dx, float* dw, float* db,
                              long long rows, int C, float eps, cudaStream_t st) {
  const int rpw = rows > (1 << 18) ? 64 : (rows > (1 << 14) ? 16 : (rows > 2048 ? 4 : 1));
  const long long warps = (rows + rpw - 1) / rpw;
  const unsigned grid = (unsigned)((warps + 7) / 8);
#define LNB(NCH, VEC)                                                                                                              \
  layernorm_bwd_reg_kernel<T, NCH, VEC><<<grid, 256, 0, st>>>((const T*)x, (const T*)dy, w, bias, gelu, (T*)dx, dw, db, rows, eps, rpw); \
  return true
  switch (C) {
    case 32: LNB(1, 1);
    case 64: LNB(1, 2);
    case 96: LNB(3, 1);
    case 128: LNB(2, 2);
    case 192: LNB(3, 2);
    case 256: LNB(4, 2);
    case 384: LNB(6, 2);
    case 512: LNB(8, 2);
    case 768: LNB(12, 2);
    default: return false;
  }
#undef LNB
}

// ================================ attention backward ================================================
// CTA per (sequence, head); Q, K, V, dO in shared memory; pass A (thread per query): lse, D, dQ;
// pass B (thread per key): dK, dV.
__device__ __forceinline__ unsigned attn_hash(unsigned long long seed, int bh, int i, int j) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)bh << 26) | ((unsigned long long)i << 13) | (unsigned long long)j);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}

template <int DH>
__global__ void __launch_bounds__(512) attention_bwd_kernel(const void* qkv, int dt, const void* dout, int do_dt, const int* cu,
                                                            int n_heads, float drop_p, AcbSeed seed_s, void* dqkv, int dq_dt,
                                                            const int* __restrict__ seq_list, const int* __restrict__ n_list) {
  if (seq_list && (int)blockIdx.x >= *n_list) return;  // list mode (long sequences of the packed plan): grid = host-side upper bound
  const unsigned long long seed = seed_s.get();
  extern __shared__ float sm[];
  const int b = seq_list ? seq_list[blockIdx.x] : (int)blockIdx.x, h = blockIdx.y;
  const int t0 = cu[b], n = cu[b + 1] - t0;
  const int D = n_heads * DH;
  float* Qs = sm;
  float* Ks = Qs + (size_t)n * DH;
  float* Vs = Ks + (size_t)n * DH;
  float* Os = Vs + (size_t)n * DH;   // dO
  float* lse = Os + (size_t)n * DH;  // [n]
  float* Dr = lse + n;               // [n]
  const float scale = rsqrtf((float)DH);
  const float drop_inv = 1.0f / (1.0f - drop_p);
  const unsigned drop_thr = (unsigned)(drop_p * 4294967296.0);
  const int bh = b * n_heads + h;
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    const int r = i / DH, c = i - r * DH;
    const long long row = (long long)(t0 + r) * 3 * D;
    Qs[i] = ld_any(qkv, row + h * DH + c, dt) * scale;
    Ks[i] = ld_any(qkv, row + D + h * DH + c, dt);
    Vs[i] = ld_any(qkv, row + 2 * D + h * DH + c, dt);
    Os[i] = ld_any(dout, (long long)(t0 + r) * D + h * DH + c, do_dt);
  }
  __syncthreads();
  // the dot products and rank-1 updates over the DH = 16 head channels run on the packed fp32 pipe (FFMA2: two channels per
  // instruction); a row of K / V / Q / dO is four 16-byte shared-memory loads
  auto dot16 = [](const float2 (&a)[DH / 2], const float* __restrict__ row) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 v = r4[c];
      acc = __ffma2_rn(a[2 * c], make_float2(v.x, v.y), acc);
      acc = __ffma2_rn(a[2 * c + 1], make_float2(v.z, v.w), acc);
    }
    return acc.x + acc.y;
  };
  auto axpy16 = [](float2 (&y)[DH / 2], float a, const float* __restrict__ row) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float2 a2 = make_float2(a, a);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 v = r4[c];
      y[2 * c] = __ffma2_rn(a2, make_float2(v.x, v.y), y[2 * c]);
      y[2 * c + 1] = __ffma2_rn(a2, make_float2(v.z, v.w), y[2 * c + 1]);
    }
  };
  auto load16 = [](float2 (&a)[DH / 2], const float* __restrict__ row) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const float4 v = r4[c];
      a[2 * c] = make_float2(v.x, v.y);
      a[2 * c + 1] = make_float2(v.z, v.w);
    }
  };
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    float2 q[DH / 2], go[DH / 2];
    load16(q, Qs + r * DH);
    load16(go, Os + r * DH);
    float m = -INFINITY;
    for (int j = 0; j < n; ++j) m = fmaxf(m, dot16(q, Ks + j * DH));
    float l = 0.0f, dsum = 0.0f;
    for (int j = 0; j < n; ++j) {
      const float s = dot16(q, Ks + j * DH);
      float dp = dot16(go, Vs + j * DH);
      const float e = expf(s - m);
      l += e;
      if (drop_p > 0.0f) dp = attn_hash(seed, bh, r, j) >= drop_thr ? dp * drop_inv : 0.0f;
      dsum += e * dp;
    }
    const float L = m + logf(l);
    const float Dv = dsum / l;  // sum_j P_ij dP_ij
    lse[r] = L;
    Dr[r] = Dv;
    float2 dq[DH / 2];
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) dq[c] = make_float2(0.0f, 0.0f);
    for (int j = 0; j < n; ++j) {
      const float s = dot16(q, Ks + j * DH);
      float dp = dot16(go, Vs + j * DH);
      if (drop_p > 0.0f) dp = attn_hash(seed, bh, r, j) >= drop_thr ? dp * drop_inv : 0.0f;
      const float ds = expf(s - L) * (dp - Dv);
      axpy16(dq, ds, Ks + j * DH);
    }
    const long long row = (long long)(t0 + r) * 3 * D + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) {
      st_any(dqkv, row + 2 * c, dq_dt, dq[c].x * scale);
      st_any(dqkv, row + 2 * c + 1, dq_dt, dq[c].y * scale);
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float2 k[DH / 2], v[DH / 2], dk[DH / 2], dv[DH / 2];
    load16(k, Ks + j * DH);
    load16(v, Vs + j * DH);
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) { dk[c] = make_float2(0.0f, 0.0f); dv[c] = make_float2(0.0f, 0.0f); }
    for (int r = 0; r < n; ++r) {
      const float s = dot16(k, Qs + r * DH);
      float dp = dot16(v, Os + r * DH);
      const float pr = expf(s - lse[r]);
      float pd = pr;
      if (drop_p > 0.0f) {
        const bool keep = attn_hash(seed, bh, r, j) >= drop_thr;
        dp = keep ? dp * drop_inv : 0.0f;
        pd = keep ? pr * drop_inv : 0.0f;
      }
      const float ds = pr * (dp - Dr[r]);
      axpy16(dv, pd, Os + r * DH);
      axpy16(dk, ds, Qs + r * DH);
    }
    const long long row = (long long)(t0 + j) * 3 * D + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) {
      st_any(dqkv, row + D + 2 * c, dq_dt, dk[c].x);  // Qs already carries the 1/sqrt(dh) factor
      st_any(dqkv, row + D + 2 * c + 1, dq_dt, dk[c].y);
      st_any(dqkv, row + 2 * D + 2 * c, dq_dt, dv[c].x);
      st_any(dqkv, row + 2 * D + 2 * c + 1, dq_dt, dv[c].y);
    }
  }
}

// ================================ photometry embedding backward =====================================
// grads[0:7D] d in_proj.weight (D,7) | [7D:8D] d in_proj.bias | [8D] dw0 | [8D+1] db0 | [8D+2 : 9D+1] dw (D-1)
// | [9D+1 : 10D] db (D-1) | [10D : 11D] d cls_tok
__global__ void __launch_bounds__(128) photo_embed_bwd_kernel(const float* x, const int* src, int T, int D, const void* dh, int dh_dt,
                                                              const float* w, const float* bb, int tok_per_block, float te_drop_p,
                                                              AcbSeed te_seed_s, float* grads) {
  const unsigned long long te_seed = te_seed_s.get();
  const int c = threadIdx.x;  // channel
  const float drop_inv = 1.0f / (1.0f - te_drop_p);
  const unsigned drop_thr = (unsigned)(te_drop_p * 4294967296.0);
  const int t0 = blockIdx.x * tok_per_block, t1 = min(T, t0 + tok_per_block);
  if (c >= D) return;
  float gw[7] = {}, gb = 0.f, g0 = 0.f, g1 = 0.f, gc = 0.f;
  for (int t = t0; t < t1; ++t) {
    const float g = ld_any(dh, (long long)t * D + c, dh_dt);
    const int s = src[t];
    if (s <= ACB_SRC_DEAD) continue;  // capacity row, no token
    if (s < 0) { gc += g; continue; }
    const float* xr = x + (long long)s * 7;
#pragma unroll
    for (int j = 0; j < 7; ++j) gw[j] = fmaf(g, xr[j], gw[j]);
    gb += g;
    const float tt = xr[0];
    float gt = g;  // gradient reaching the Time2Vec term (dropout mask regenerated from the hash)
    if (te_drop_p > 0.0f) gt = (te_hash(te_seed, t, c) < drop_thr) ? 0.0f : g * drop_inv;
    if (c == 0) { g0 += gt * tt; g1 += gt; }
    else { const float cs = cosf(tt * w[c - 1] + bb[c - 1]); g0 += gt * cs * tt; g1 += gt * cs; }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) atomicAdd(grads + c * 7 + j, gw[j]);
  atomicAdd(grads + 7 * D + c, gb);
  if (c == 0) { atomicAdd(grads + 8 * D, g0); atomicAdd(grads + 8 * D + 1, g1); }
  else { atomicAdd(grads + 8 * D + 2 + (c - 1), g0); atomicAdd(grads + 9 * D + 1 + (c - 1), g1); }
  atomicAdd(grads + 10 * D + c, gc);
}

__global__ void scatter_cls_kernel(const float* dcls, const int* cu, int B, int D, void* dh, int dh_dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D) return;
  const int b = (int)(i / D), c = (int)(i % D);
  st_any(dh, (long long)cu[b] * D + c, dh_dt, dcls[i]);
}

// ================================ depthwise 7x7 (no LN): fwd / bwd-data / bwd-weight ==================
// flip = 0: y = conv(x, w) + bias ; flip = 1: y = conv(x, flipped w)  (gradient w.r.t. the input)
__global__ void __launch_bounds__(256) dwconv7_kernel(const void* x, int x_dt, const float* w, const float* bias, int flip, void* y,
                                                      int y_dt, int B, int H, int W, int C) {
  const long long tot = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int ox = (int)(t % W);
    t /= W;
    const int oy = (int)(t % H);
    const long long b = t / H;
    float acc = bias ? bias[c] : 0.0f;
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = oy + ky - 3;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 7; ++kx) {
        const int ix = ox + kx - 3;
        if (ix < 0 || ix >= W) continue;
        const float wv = flip ? w[c * 49 + (6 - ky) * 7 + (6 - kx)] : w[c * 49 + ky * 7 + kx];
        acc = fmaf(ld_any(x, ((b * H + iy) * W + ix) * C + c, x_dt), wv, acc);
      }
    }
    st_any(y, i, y_dt, acc);
  }
}

// dw[c,ky,kx] += sum_{b,y,x} dy[b,y,x,c] * x[b,y+ky-3,x+kx-3,c]; db[c] += sum dy.  CTA per image chunk, thread per channel.
__global__ void __launch_bounds__(256) dwconv7_wgrad_kernel(const void* x, int x_dt, const void* dy, int dy_dt, int B, int H, int W, int C,
                                                            int img_per_block, float* dw, float* db) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc[49];
#pragma unroll
  for (int j = 0; j < 49; ++j) acc[j] = 0.0f;
  float sb = 0.0f;
  const int b0 = blockIdx.x * img_per_block, b1 = min(B, b0 + img_per_block);
  for (int b = b0; b < b1; ++b) {
    for (int oy = 0; oy < H; ++oy) {
      for (int ox = 0; ox < W; ++ox) {
        const float g = ld_any(dy, (((long long)b * H + oy) * W + ox) * C + c, dy_dt);
        sb += g;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const int iy = oy + ky - 3;
          if (iy < 0 || iy >= H) continue;
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const int ix = ox + kx - 3;
            if (ix < 0 || ix >= W) continue;
            acc[ky * 7 + kx] = fmaf(g, ld_any(x, (((long long)b * H + iy) * W + ix) * C + c, x_dt), acc[ky * 7 + kx]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 49; ++j) atomicAdd(dw + c * 49 + j, acc[j]);
  atomicAdd(db + c, sb);
}

// 1x1 maps (ConvNeXt stage 3 at 63x63 cutouts): only the centre tap ever sees data: dw[c, 3, 3] = sum_b x[b,c] dy[b,c], db[c] = sum_b dy[b,c].
// thread = channel (coalesced), blockIdx.y strides over image chunks; two atomics per thread instead of 50.
__global__ void __launch_bounds__(256) dwconv7_wgrad_1x1_kernel(const void* __restrict__ x, int x_dt, const void* __restrict__ dy, int dy_dt, int B,
                                                                int C, int img_per_block, float* __restrict__ dw, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int b0 = blockIdx.y * img_per_block, b1 = min(B, b0 + img_per_block);
  float acc = 0.0f, sb = 0.0f;
  for (int b = b0; b < b1; ++b) {
    const float g = ld_any(dy, (long long)b * C + c, dy_dt);
    sb += g;
    acc = fmaf(g, ld_any(x, (long long)b * C + c, x_dt), acc);
  }
  atomicAdd(dw + c * 49 + 24, acc);
  atomicAdd(db + c, sb);
}

// Row-register variant for the square ConvNeXt maps (W = H in {15, 7, 3}): thread = (channel, row phase); the dy row
// and one x row live in registers, so every x value loaded feeds up to 7 taps (loads : FMAs = 1 : 7 instead of 1 : 1);
// partial sums are merged in shared memory before the global atomics.
template <typename T, int W>
__global__ void __launch_bounds__(512) dwconv7_wgrad_w_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int C, int cc,
                                                              int rowsplit, int img_per_block, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float red[];  // [cc][50]
  for (int i = threadIdx.x; i < cc * 50; i += blockDim.x) red[i] = 0.0f;
  __syncthreads();
  const int cl = threadIdx.x % cc, part = threadIdx.x / cc;
  const int c = blockIdx.y * cc + cl;
  if (c < C) {
    float acc[49];
#pragma unroll
    for (int j = 0; j < 49; ++j) acc[j] = 0.0f;
    float sb = 0.0f;
    const int b0 = blockIdx.x * img_per_block, b1 = min(B, b0 + img_per_block);
    for (int b = b0; b < b1; ++b) {
      const T* xb = x + (long long)b * W * W * C + c;
      const T* gb = dy + (long long)b * W * W * C + c;
      for (int oy = part; oy < W; oy += rowsplit) {
        float g[W];
#pragma unroll
        for (int ox = 0; ox < W; ++ox) {
          g[ox] = to_f<T>(gb[(long long)(oy * W + ox) * C]);
          sb += g[ox];
        }
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const int iy = oy + ky - 3;
          if (iy < 0 || iy >= W) continue;
          float xr[W];
#pragma unroll
          for (int ix = 0; ix < W; ++ix) xr[ix] = to_f<T>(xb[(long long)(iy * W + ix) * C]);
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
            for (int ox = 0; ox < W; ++ox) {
              const int ix = ox + kx - 3;
              if (ix >= 0 && ix < W) acc[ky * 7 + kx] = fmaf(g[ox], xr[ix], acc[ky * 7 + kx]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 49; ++j) atomicAdd(&red[cl * 50 + j], acc[j]);
    atomicAdd(&red[cl * 50 + 49], sb);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cc * 50; i += blockDim.x) {
    const int l = i / 50, j = i - l * 50, ch = blockIdx.y * cc + l;
    if (ch >= C) continue;
    if (j < 49) atomicAdd(dw + ch * 49 + j, red[i]);
    else atomicAdd(db + ch, red[i]);
  }
}

template <typename T>
static bool launch_dwconv7_wgrad_w(const void* x, const void* dy, int B, int H, int W, int C, float* dw, float* db, cudaStream_t st) {
  if (H != W || !(W == 15 || W == 7 || W == 3)) return false;
  const int cc = C >= 128 ? 128 : ((C + 31) / 32) * 32;
  const int rowsplit = min(W, 512 / cc);
  int ipb = 1;
  while (ipb < 16 && (long long)cdiv(B, ipb * 2) * cdiv(C, cc) >= 2 * 148) ipb *= 2;
  const dim3 grid(cdiv(B, ipb), cdiv(C, cc));
  const size_t smem = (size_t)cc * 50 * 4;
  if (W == 15) dwconv7_wgrad_w_kernel<T, 15><<<grid, cc * rowsplit, smem, st>>>((const T*)x, (const T*)dy, B, C, cc, rowsplit, ipb, dw, db);
  else if (W == 7) dwconv7_wgrad_w_kernel<T, 7><<<grid, cc * rowsplit, smem, st>>>((const T*)x, (const T*)dy, B, C, cc, rowsplit, ipb, dw, db);
  else dwconv7_wgrad_w_kernel<T, 3><<<grid, cc * rowsplit, smem, st>>>((const T*)x, (const T*)dy, B, C, cc, rowsplit, ipb, dw, db);
  return true;
}

// 2x2/stride-2 patch gather (fwd) and its adjoint (bwd: rows/cols dropped by the floor get zero)
__global__ void patch2_kernel(void* x, int x_dt, void* p, int p_dt, int B, int H, int W, int C, int adjoint) {
  const int Ho = H / 2, Wo = W / 2;
  const long long tot = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const long long b = t / H;
    const bool inside = iy < 2 * Ho && ix < 2 * Wo;
    const long long pi = ((b * Ho + (iy >> 1)) * Wo + (ix >> 1)) * (4LL * C) + ((iy & 1) * 2 + (ix & 1)) * C + c;
    if (!adjoint) {
      if (inside) st_any(p, pi, p_dt, ld_any(x, i, x_dt));
    } else {
      st_any(x, i, x_dt, inside ? ld_any(p, pi, p_dt) : 0.0f);  // here x = dx (out), p = dpatches (in)
    }
  }
}

// mean over HW: y[b,c] = mean_i x[b,i,c] (fwd) ; dx[b,i,c] = dy[b,c]/HW (bwd)
__global__ void gap_kernel(const void* x, int x_dt, float* y, int B, int HW, int C, int bwd, void* dx, int dx_dt) {
  const long long tot = (long long)B * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long b = i / C;
    if (!bwd) {
      float s = 0.0f;
      for (int k = 0; k < HW; ++k) s += ld_any(x, (b * HW + k) * C + c, x_dt);
      y[i] = s / (float)HW;
    } else {
      const float g = y[i] / (float)HW;
      for (int k = 0; k < HW; ++k) st_any(dx, (b * HW + k) * C + c, dx_dt, g);
    }
  }
}

// ================================ pooling backward ==================================================
// window = 4: MaxPool1d(4) over L of [B,L,C] (grad to the first max of each window, zero elsewhere incl. the floor tail)
// window = 0: global max over L
__global__ void maxpool_bwd_kernel(const void* x, int x_dt, const void* dy, int dy_dt, void* dx, int dx_dt, int B, int L, int C, int window) {
  const int Lo = window ? L / window : 1;
  const int win = window ? window : L;
  const long long tot = (long long)B * Lo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int lo = (int)(t % Lo);
    const long long b = t / Lo;
    const long long base = (b * L + (long long)lo * win) * C + c;
    float m = -INFINITY;
    int arg = 0;
    for (int k = 0; k < win; ++k) {
      const float v = ld_any(x, base + (long long)k * C, x_dt);
      if (v > m) { m = v; arg = k; }
    }
    const float g = ld_any(dy, i, dy_dt);
    for (int k = 0; k < win; ++k) st_any(dx, base + (long long)k * C, dx_dt, k == arg ? g : 0.0f);
    if (window && lo == Lo - 1) {
      for (int l = Lo * win; l < L; ++l) st_any(dx, (b * L + l) * C + c, dx_dt, 0.0f);
    }
  }
}

// bf16, window 4, C % 8 == 0: a thread owns 8 channels of one window -- four 16-byte loads of x, one of dy, four 16-byte stores
__global__ void __launch_bounds__(256) maxpool4_bwd_bf16x8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                                                  long long B, int L, int C8) {
  const int Lo = L / 4;
  const long long tot = B * Lo * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    const long long t = i / C8;
    const int lo = (int)(t % Lo);
    const long long b = t / Lo;
    const long long base = (b * L + (long long)lo * 4) * C8 + c;
    uint4 xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xv[k] = __ldcs(x + base + (long long)k * C8);
    const uint4 g = __ldcs(dy + i);
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g);
    uint4 o[4];
    uint32_t* ow[4] = {reinterpret_cast<uint32_t*>(&o[0]), reinterpret_cast<uint32_t*>(&o[1]), reinterpret_cast<uint32_t*>(&o[2]),
                       reinterpret_cast<uint32_t*>(&o[3])};
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // two channels per 32-bit word
      float m0 = -INFINITY, m1 = -INFINITY;
      int a0 = 0, a1 = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&xv[k]) + j));
        if (v.x > m0) { m0 = v.x; a0 = k; }
        if (v.y > m1) { m1 = v.y; a1 = k; }
      }
      const uint32_t glo = gw[j] & 0xffffu, ghi = gw[j] & 0xffff0000u;
#pragma unroll
      for (int k = 0; k < 4; ++k) ow[k][j] = (a0 == k ? glo : 0u) | (a1 == k ? ghi : 0u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dx[base + (long long)k * C8] = o[k];
    if (lo == Lo - 1)
      for (int l = Lo * 4; l < L; ++l) dx[(b * L + l) * C8 + c] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ================================ MoE / L2 norm / losses / dropout ==================================
__global__ void moe_combine_bwd_kernel(const float* gate, const float* eo, const float* dout, float* dgate, float* deo, int B, int E, int C) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  int i0 = 0;
  float v0 = gate[(long long)r * E];
  for (int e = 1; e < E; ++e) { const float v = gate[(long long)r * E + e]; if (v > v0) { v0 = v; i0 = e; } }
  int i1 = -1;
  float v1 = -INFINITY;
  for (int e = 0; e < E; ++e) { if (e == i0) continue; const float v = gate[(long long)r * E + e]; if (v > v1) { v1 = v; i1 = e; } }
  for (int e = 0; e < E; ++e) {
    const bool sel = (e == i0) || (e == i1);
    const float wv = e == i0 ? v0 : v1;
    float dg = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float go = dout[(long long)r * C + c];
      const long long idx = ((long long)r * E + e) * C + c;
      if (sel) dg = fmaf(go, eo[idx], dg);
      deo[idx] = sel ? wv * go : 0.0f;
    }
    dgate[(long long)r * E + e] = sel ? dg : 0.0f;
  }
}

// y = x / ||x||_2 per row (fwd) ; dx = (dy - y * (y.dy)) / ||x||  (bwd)
__global__ void l2norm_kernel(const float* x, const float* dy, float* out, int rows, int C, int bwd) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float ss = 0.0f;
  for (int c = lane; c < C; c += 32) { const float v = x[(long long)r * C + c]; ss += v * v; }
  const float nrm = sqrtf(warp_sum(ss));
  if (!bwd) {
    for (int c = lane; c < C; c += 32) out[(long long)r * C + c] = x[(long long)r * C + c] / nrm;
  } else {
    float dot = 0.0f;
    for (int c = lane; c < C; c += 32) dot += (x[(long long)r * C + c] / nrm) * dy[(long long)r * C + c];
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) out[(long long)r * C + c] = (dy[(long long)r * C + c] - (x[(long long)r * C + c] / nrm) * dot) / nrm;
  }
}

// focal loss (gamma, mean reduction) with integer labels, or soft-target cross entropy (targets [B,C], mean):
// loss_out[0] += per-row loss / B ; dlogits = d(mean loss)/d logits
__global__ void loss_kernel(const float* logits, const long long* labels, const float* soft, float gamma, int B, int C, float* loss_out,
                            float* dlogits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float lr = 0.0f;
  if (r < B) {
    const float* z = logits + (long long)r * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, z[c]);
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += expf(z[c] - m);
    const float lz = m + logf(s);
    if (labels) {
      const int y = (int)labels[r];
      const float logp = z[y] - lz, p = expf(logp);
      const float om = 1.0f - p;
      const float fw = powf(om, gamma);
      lr = -fw * logp;
      // d/dlogp [-(1-p)^g logp] = g (1-p)^(g-1) p logp - (1-p)^g
      const float dldlogp = gamma * powf(om, gamma - 1.0f) * p * logp - fw;
      for (int c = 0; c < C; ++c) {
        const float pc = expf(z[c] - lz);
        dlogits[(long long)r * C + c] = dldlogp * ((c == y ? 1.0f : 0.0f) - pc) / (float)B;
      }
    } else {
      const float* t = soft + (long long)r * C;
      float ts = 0.0f;
      for (int c = 0; c < C; ++c) { lr -= t[c] * (z[c] - lz); ts += t[c]; }
      for (int c = 0; c < C; ++c) dlogits[(long long)r * C + c] = (expf(z[c] - lz) * ts - t[c]) / (float)B;
    }
    lr /= (float)B;
  }
  lr = warp_sum(lr);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss_out, lr);
}

__device__ __forceinline__ unsigned hash32(unsigned long long v) {
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}
// y = keep ? x / (1-p) : 0, keep decided by a counter-based hash of (seed, index); the same call with dy gives dx
// (one 64-bit hash decides the element pair (2j, 2j+1): low word / high word -- the scalar and the vector kernel draw the SAME mask,
// so a forward through one and a backward through the other stay consistent)
__device__ __forceinline__ unsigned long long hash64(unsigned long long v) {
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return v;
}
__global__ void dropout_kernel(const void* x, int x_dt, void* y, int y_dt, float p, AcbSeed seed_s, long long n) {
  const unsigned long long seed = seed_s.get() * 0x9E3779B97F4A7C15ULL;
  const float inv = 1.0f / (1.0f - p);
  const unsigned thr = (unsigned)(p * 4294967296.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h = hash64(seed + (unsigned long long)(i >> 1));
    const bool keep = (unsigned)((i & 1) ? (h >> 32) : h) >= thr;
    st_any(y, i, y_dt, keep ? ld_any(x, i, x_dt) * inv : 0.0f);
  }
}
// bf16 in and out, n % 8 == 0: 16 bytes per thread and step
__global__ void __launch_bounds__(256) dropout_bf16x8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, float p, AcbSeed seed_s, long long n8) {
  const unsigned long long seed = seed_s.get() * 0x9E3779B97F4A7C15ULL;
  const float inv = 1.0f / (1.0f - p);
  const unsigned thr = (unsigned)(p * 4294967296.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 v = __ldcs(x + i);
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned long long h = hash64(seed + (unsigned long long)(i * 4 + j));
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
      const __nv_bfloat162 o = __floats2bfloat162_rn((unsigned)h >= thr ? f.x * inv : 0.0f, (unsigned)(h >> 32) >= thr ? f.y * inv : 0.0f);
      w[j] = *reinterpret_cast<const uint32_t*>(&o);
    }
    y[i] = v;
  }
}

// sum of squares (grad-norm) : out[0] += sum x^2
__global__ void sumsq_kernel(const float* x, long long n, float* out) {
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

}  // namespace

int acb_dwconv7_fast(const void* x, int dtype, const float* w, const float* bias, int flip, void* y, int B, int H, int W, int C, cudaStream_t st);

#define LAUNCHED(n)     \
  ACB_LAUNCH_CHECK();   \
  acb_count_launch(n);  \
  return ACB_OK

extern "C" {

int acb_gemm_ex(const void* A, int a_dtype, const void* B, int b_dtype, float* C, int M, int N, int K, long long sam, long long sak,
                long long sbn, long long sbk, int ldc, int convT, int conv_L, int conv_Cin, int conv_pad, int splits, int accumulate,
                void* stream) {
  ACB_CHECK(A && B && C && M > 0 && N > 0 && K > 0 && splits >= 1, "acb_gemm_ex: bad arguments");
  GemmExArgs p{A, B, C, a_dtype, b_dtype, M, N, K, sam, sak, sbn, sbk, ldc, convT, conv_L, conv_Cin, conv_pad, 0, 0};
  int chunk = (K + splits - 1) / splits;
  chunk = ((chunk + 15) / 16) * 16;
  splits = (K + chunk - 1) / chunk;
  p.k_chunk = chunk;
  p.atomic = (splits > 1 || accumulate) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (splits > 1 && !accumulate) ACB_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, st));
  dim3 grid(cdiv(M, 64), cdiv(N, 64), splits);
  ACB_CHECK(grid.y <= 65535 && grid.z <= 65535, "acb_gemm_ex: grid too large");
  gemm_ex_kernel<<<grid, 256, 0, st>>>(p);
  LAUNCHED(1);
}

int acb_colsum(const void* a, int a_dtype, const void* b, int b_dtype, long long M, int N, long long ld, float* out, int accumulate,
               void* stream) {
  if (ld <= 0) ld = N;
  ACB_CHECK(a && out && M > 0 && N > 0, "acb_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) ACB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * 4, st));
  if (a_dtype == ACB_BF16 && (!b || b_dtype == ACB_BF16) && N % 8 == 0 && N <= 2048 && ld % 8 == 0 && M >= 512 &&
      (((uintptr_t)a | (uintptr_t)b) & 15) == 0) {
    // ~4 CTAs per SM, at least 8 rows per row lane
    const int lanes = 256 / (N / 8);
    const int rpbv = (int)std::max<long long>(8LL * lanes, cdiv(M, 148 * 4));
    colsum_bf16x8_kernel<<<(unsigned)cdiv(M, rpbv), 256, 0, st>>>((const uint4*)a, (const uint4*)b, M, N / 8, ld / 8, rpbv, out);
    LAUNCHED(1);
  }
  const int rpb = 1024;
  dim3 grid(cdiv(N, 32), cdiv(M, rpb));
  colsum_kernel<<<grid, 256, 0, st>>>(a, a_dtype, b, b_dtype, M, N, ld, rpb, out);
  LAUNCHED(1);
}

int acb_act_fwd(const void* x, int x_dtype, void* y, int y_dtype, int act, long long n, void* stream) {
  ACB_CHECK(x && y && n >= 0, "acb_act_fwd: bad arguments");
  if (n == 0) return ACB_OK;
  act_fwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, y_dtype, act, n);
  LAUNCHED(1);
}

int acb_act_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, void* dx, int dx_dtype, int act, long long n, void* stream) {
  ACB_CHECK(dy && x && dx && n >= 0, "acb_act_bwd: bad arguments");
  if (n == 0) return ACB_OK;
  if ((act == ACB_ACT_RELU || act == ACB_ACT_GELU) && dy_dtype == ACB_BF16 && x_dtype == ACB_BF16 && dx_dtype == ACB_BF16 && n % 8 == 0 &&
      (((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx) & 15) == 0) {
    act_bwd_bf16x8_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)x, (uint4*)dx, act, n / 8);
    LAUNCHED(1);
  }
  act_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dy, dy_dtype, x, x_dtype, dx, dx_dtype, act, n);
  LAUNCHED(1);
}

int acb_ew(const void* a, int a_dtype, const void* b, int b_dtype, const float* g, void* y, int y_dtype, int op, int C, float s0, float s1,
           long long n, void* stream) {
  ACB_CHECK(a && y && n >= 0 && C > 0, "acb_ew: bad arguments");
  if (n == 0) return ACB_OK;
  const bool two = op == 0 || op == 1 || op == 2 || op == 4, col = op == 2 || op == 3;
  if ((two || op == 3) && (b != nullptr) == two && (!two || b_dtype == ACB_BF16) && a_dtype == ACB_BF16 && y_dtype == ACB_BF16 && n % 8 == 0 &&
      (!col || (g && C % 8 == 0 && ((uintptr_t)g & 15) == 0)) && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 15) == 0) {
    ew_bf16x8_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, (const uint4*)b, g, (uint4*)y, op, C, s0, s1, n / 8);
    LAUNCHED(1);
  }
  ew_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, a_dtype, b, b_dtype, g, y, y_dtype, op, C, s0, s1, n);
  LAUNCHED(1);
}

int acb_copy2d(const void* src, int s_dtype, long long lds, void* dst, int d_dtype, long long ldd, long long rows, int cols, void* stream) {
  ACB_CHECK(src && dst && rows >= 0 && cols > 0, "acb_copy2d: bad arguments");
  if (rows == 0) return ACB_OK;
  copy2d_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(src, s_dtype, lds, dst, d_dtype, ldd, rows, cols);
  LAUNCHED(1);
}

int acb_unpack_conv_wgrad(const float* G, float* dW, int Cout, int Cin, int k, void* stream) {
  ACB_CHECK(G && dW && Cout > 0 && Cin > 0 && k > 0, "acb_unpack_conv_wgrad: bad arguments");
  unpack_conv_wgrad_kernel<<<grid_for((long long)Cout * Cin * k), 256, 0, (cudaStream_t)stream>>>(G, dW, Cout, Cin, k);
  LAUNCHED(1);
}

int acb_pack_conv_dgrad_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int k, void* stream) {
  ACB_CHECK(w && out && Cout > 0 && Cin > 0 && k > 0, "acb_pack_conv_dgrad_weight: bad arguments");
  pack_conv_dgrad_kernel<<<grid_for((long long)Cout * Cin * k), 256, 0, (cudaStream_t)stream>>>(w, out, out_dtype, Cout, Cin, k);
  LAUNCHED(1);
}

int acb_gather_cols(const float* X, int ldx, const int* cols, int n, float* Y, long long rows, void* stream) {
  ACB_CHECK(X && cols && Y && n > 0 && rows >= 0, "acb_gather_cols: bad arguments");
  if (rows == 0) return ACB_OK;
  gather_cols_kernel<<<grid_for(rows * n), 256, 0, (cudaStream_t)stream>>>(X, ldx, cols, n, Y, rows);
  LAUNCHED(1);
}

int acb_layernorm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* w, const float* b, int post_act, void* dx,
                      int dx_dtype, float* dw, float* db, long long rows, int C, float eps, void* stream) {
  ACB_CHECK(x && dy && w && dx && dw && db && rows >= 0 && C > 0 && C <= 6000, "acb_layernorm_bwd: bad arguments");
  ACB_CHECK(post_act == ACB_ACT_NONE || (post_act == ACB_ACT_GELU && b), "acb_layernorm_bwd: post_act must be none or GELU (with the LayerNorm bias)");
  const int gelu = post_act == ACB_ACT_GELU;
  if (rows == 0) return ACB_OK;
  if (x_dtype == dy_dtype && x_dtype == dx_dtype) {
    const bool done = x_dtype == ACB_F32 ? launch_ln_bwd_reg<float>(x, dy, w, b, gelu, dx, dw, db, rows, C, eps, (cudaStream_t)stream)
                                         : launch_ln_bwd_reg<bf16>(x, dy, w, b, gelu, dx, dw, db, rows, C, eps, (cudaStream_t)stream);
    if (done) { LAUNCHED(1); }
    if (x_dtype == ACB_BF16 && rows >= 1024 && launch_ln_bwd_wide(x, dy, w, b, gelu, dx, dw, db, rows, C, eps, (cudaStream_t)stream)) { LAUNCHED(1); }
  }
  const int rpw = rows > (1 << 16) ? 16 : 1;
  const long long warps = (rows + rpw - 1) / rpw;
  layernorm_bwd_kernel<<<(unsigned)((warps + 7) / 8), 256, (size_t)2 * C * 4, (cudaStream_t)stream>>>(x, x_dtype, dy, dy_dtype, w, b, gelu, dx,
                                                                                                    dx_dtype, dw, db, rows, C, eps, rpw);
  LAUNCHED(1);
}

int acb_attention_varlen_bwd(const void* qkv, int dtype, const void* dout, int dout_dtype, const int* cu_seqlens, int B, int n_heads, int dh,
                             int max_seqlen, float drop_p, long long seed, void* dqkv, int dqkv_dtype, void* stream) {
  ACB_CHECK(qkv && dout && cu_seqlens && dqkv && B > 0 && dh == 16, "acb_attention_varlen_bwd: bad arguments (dh must be 16)");
  const size_t smem = ((size_t)max_seqlen * dh * 4 + 2 * (size_t)max_seqlen) * 4;
  ACB_CHECK(smem <= 200 * 1024, "acb_attention_varlen_bwd: max_seqlen %d too long", max_seqlen);
  auto k = attention_bwd_kernel<16>;
  ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int threads = 128;  // measured: 288-thread CTAs (one round for 258-token sequences) are 25 % slower overall
  k<<<dim3(B, n_heads), threads, smem, (cudaStream_t)stream>>>(qkv, dtype, dout, dout_dtype, cu_seqlens, n_heads, drop_p, acb_seed(seed), dqkv, dqkv_dtype,
                                                                nullptr, nullptr);
  LAUNCHED(1);
}

}  // extern "C"

// the same kernel over a device-side list of sequences (attention_packed.cu: sequences longer than one tile)
int acb_attention_bwd_long(const void* qkv, const void* dout, const int* cu_seqlens, const int* long_list, const int* n_long_dev, int grid_x,
                           int n_heads, int max_seqlen, float drop_p, long long seed, void* dqkv, cudaStream_t st) {
  const size_t smem = ((size_t)max_seqlen * 16 * 4 + 2 * (size_t)max_seqlen) * 4;
  ACB_CHECK(smem <= 200 * 1024, "acb_attention_packed_bwd: max_seqlen %d too long", max_seqlen);
  auto k = attention_bwd_kernel<16>;
  ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // every sequence of the list has more than 128 tokens: one thread per token (the 128-thread default would run a second round
  // with a handful of active threads)
  const int threads = max_seqlen <= 512 ? ((max_seqlen + 31) / 32) * 32 : 512;
  k<<<dim3(grid_x, n_heads), threads, smem, st>>>(qkv, ACB_BF16, dout, ACB_BF16, cu_seqlens, n_heads, drop_p, acb_seed(seed), dqkv, ACB_BF16, long_list, n_long_dev);
  LAUNCHED(1);
}

extern "C" {

int acb_photo_embed_bwd(const float* x, const int* src_idx, int T, int D, const void* dh, int dh_dtype, const float* w, const float* b,
                        float te_drop_p, long long te_seed, float* grads, void* stream) {
  ACB_CHECK(x && src_idx && dh && w && b && grads && T >= 0 && D > 1 && D <= 128, "acb_photo_embed_bwd: bad arguments (D <= 128)");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(grads, 0, (size_t)11 * D * 4, st));
  if (T == 0) return ACB_OK;
  const int tpb = 256;
  photo_embed_bwd_kernel<<<cdiv(T, tpb), 128, 0, st>>>(x, src_idx, T, D, dh, dh_dtype, w, b, tpb, te_drop_p, acb_seed(te_seed), grads);
  LAUNCHED(1);
}

int acb_scatter_cls(const float* dcls, const int* cu_seqlens, int B, int D, void* dh, int dh_dtype, long long total_tokens, void* stream) {
  ACB_CHECK(dcls && cu_seqlens && dh && B > 0 && D > 0, "acb_scatter_cls: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(dh, 0, (size_t)total_tokens * D * (dh_dtype == ACB_F32 ? 4 : 2), st));
  scatter_cls_kernel<<<cdiv((long long)B * D, 256), 256, 0, st>>>(dcls, cu_seqlens, B, D, dh, dh_dtype);
  LAUNCHED(1);
}

int acb_dwconv7(const void* x, int x_dtype, const float* w, const float* bias, int flip, void* y, int y_dtype, int B, int H, int W, int C,
                void* stream) {
  ACB_CHECK(x && w && y && B > 0 && H > 0 && W > 0 && C > 0, "acb_dwconv7: bad arguments");
  if (x_dtype == y_dtype) {
    const int rc = acb_dwconv7_fast(x, x_dtype, w, bias, flip, y, B, H, W, C, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  dwconv7_kernel<<<grid_for((long long)B * H * W * C), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, w, bias, flip, y, y_dtype, B, H, W, C);
  LAUNCHED(1);
}

int acb_dwconv7_wgrad(const void* x, int x_dtype, const void* dy, int dy_dtype, int B, int H, int W, int C, float* dw, float* db, int accumulate,
                      void* stream) {
  ACB_CHECK(x && dy && dw && db && B > 0 && C > 0, "acb_dwconv7_wgrad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    ACB_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * 49 * 4, st));
    ACB_CUDA(cudaMemsetAsync(db, 0, (size_t)C * 4, st));
  }
  if (H == 1 && W == 1) {
    const int ipb = 32;
    dwconv7_wgrad_1x1_kernel<<<dim3(cdiv(C, 256), cdiv(B, ipb)), 256, 0, st>>>(x, x_dtype, dy, dy_dtype, B, C, ipb, dw, db);
    LAUNCHED(1);
  }
  if (x_dtype == dy_dtype) {
    const bool done = x_dtype == ACB_F32 ? launch_dwconv7_wgrad_w<float>(x, dy, B, H, W, C, dw, db, st) : launch_dwconv7_wgrad_w<bf16>(x, dy, B, H, W, C, dw, db, st);
    if (done) { LAUNCHED(1); }
  }
  const int ipb = B > 2048 ? 8 : (B > 256 ? 2 : 1);
  const int threads = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  dwconv7_wgrad_kernel<<<dim3(cdiv(B, ipb), cdiv(C, threads)), threads, 0, st>>>(x, x_dtype, dy, dy_dtype, B, H, W, C, ipb, dw, db);
  LAUNCHED(1);
}

int acb_patch2(const void* x, int x_dtype, void* p, int p_dtype, int B, int H, int W, int C, int adjoint, void* stream) {
  ACB_CHECK(x && p && B > 0 && H >= 2 && W >= 2 && C > 0, "acb_patch2: bad arguments");
  patch2_kernel<<<grid_for((long long)B * H * W * C), 256, 0, (cudaStream_t)stream>>>(const_cast<void*>(x), x_dtype, p, p_dtype, B, H, W, C, adjoint);
  LAUNCHED(1);
}

int acb_gap(const void* x, int x_dtype, float* y, int B, int HW, int C, int bwd, void* dx, int dx_dtype, void* stream) {
  ACB_CHECK(y && B > 0 && HW > 0 && C > 0 && (bwd ? dx != nullptr : x != nullptr), "acb_gap: bad arguments");
  gap_kernel<<<grid_for((long long)B * C), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, B, HW, C, bwd, dx, dx_dtype);
  LAUNCHED(1);
}

int acb_maxpool_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype, int B, int L, int C, int window,
                    void* stream) {
  ACB_CHECK(x && dy && dx && B > 0 && L > 0 && C > 0 && (window == 0 || window == 4), "acb_maxpool_bwd: bad arguments");
  const int Lo = window ? L / window : 1;
  if (window == 4 && Lo > 0 && C % 8 == 0 && x_dtype == ACB_BF16 && dy_dtype == ACB_BF16 && dx_dtype == ACB_BF16 &&
      (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0) {
    maxpool4_bwd_bf16x8_kernel<<<grid_for((long long)B * Lo * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)dy, (uint4*)dx, B, L, C / 8);
    LAUNCHED(1);
  }
  maxpool_bwd_kernel<<<grid_for((long long)B * Lo * C), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, dy, dy_dtype, dx, dx_dtype, B, L, C, window);
  LAUNCHED(1);
}

int acb_moe_combine_bwd(const float* gate, const float* expert_out, const float* dout, float* dgate, float* dexpert_out, int B, int E, int C,
                        void* stream) {
  ACB_CHECK(gate && expert_out && dout && dgate && dexpert_out && B > 0 && E >= 2 && C > 0, "acb_moe_combine_bwd: bad arguments");
  moe_combine_bwd_kernel<<<cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(gate, expert_out, dout, dgate, dexpert_out, B, E, C);
  LAUNCHED(1);
}

int acb_l2norm(const float* x, const float* dy, float* out, int rows, int C, int bwd, void* stream) {
  ACB_CHECK(x && out && rows > 0 && C > 0 && (!bwd || dy), "acb_l2norm: bad arguments");
  l2norm_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, dy, out, rows, C, bwd);
  LAUNCHED(1);
}

int acb_loss_fwd_bwd(const float* logits, const long long* labels, const float* soft_targets, float gamma, int B, int C, float* loss_out,
                     float* dlogits, void* stream) {
  ACB_CHECK(logits && loss_out && dlogits && B > 0 && C > 0 && ((labels != nullptr) != (soft_targets != nullptr)),
            "acb_loss_fwd_bwd: pass exactly one of labels (focal) or soft_targets (cross entropy)");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(loss_out, 0, 4, st));
  loss_kernel<<<cdiv(B, 128), 128, 0, st>>>(logits, labels, soft_targets, gamma, B, C, loss_out, dlogits);
  LAUNCHED(1);
}

int acb_dropout(const void* x, int x_dtype, void* y, int y_dtype, float p, long long seed, long long n, void* stream) {
  ACB_CHECK(x && y && n >= 0 && p >= 0.0f && p < 1.0f, "acb_dropout: bad arguments");
  if (n == 0) return ACB_OK;
  if (x_dtype == ACB_BF16 && y_dtype == ACB_BF16 && n % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
    dropout_bf16x8_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, p, acb_seed(seed), n / 8);
    LAUNCHED(1);
  }
  dropout_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, y_dtype, p, acb_seed(seed), n);
  LAUNCHED(1);
}

int acb_sumsq(const float* x, long long n, float* out, int accumulate, void* stream) {
  ACB_CHECK(x && out && n >= 0, "acb_sumsq: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) ACB_CUDA(cudaMemsetAsync(out, 0, 4, st));
  if (n == 0) return ACB_OK;
  sumsq_kernel<<<grid_for(n), 256, 0, st>>>(x, n, out);
  LAUNCHED(1);
}

}  // extern "C"
