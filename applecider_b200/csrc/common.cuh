// Common device/host helpers for the applecider_b200 sm_100a kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/applecider_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI: every entry point returns 0 or a negative code; message via acb_last_error)
// ---------------------------------------------------------------------------------------------
void acb_set_error(const char* fmt, ...);

#define ACB_CHECK(cond, ...)            \
  do {                                  \
    if (!(cond)) {                      \
      acb_set_error(__VA_ARGS__);       \
      return ACB_ERR_INVALID;           \
    }                                   \
  } while (0)

#define ACB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      acb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ACB_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define ACB_LAUNCH_CHECK() ACB_CUDA(cudaGetLastError())

void acb_count_launch(int n = 1);  // product-side launch counter (bench.py "gpu_launches")

// src_idx entries <= ACB_SRC_DEAD mark capacity rows past the last packed token (acb_photo_compact memsets them to 0x80808080)
#define ACB_SRC_DEAD (-0x40000000)

// Seed of a counter-hash RNG stream as the kernels receive it: the host value plus an optional device-resident epoch
// (acb_set_seed_epoch_ptr).  A CUDA graph bakes `base` in; bumping *epoch inside the graph gives every replay fresh masks,
// and forward / backward kernels of one replay still agree because they read the same epoch.
struct AcbSeed {
  unsigned long long base;
  const unsigned long long* epoch;
  __device__ __forceinline__ unsigned long long get() const { return base + (epoch ? *epoch * 0xD1342543DE82EF95ULL : 0ULL); }
};
const unsigned long long* acb_seed_epoch_ptr();
static inline AcbSeed acb_seed(long long s) { return AcbSeed{(unsigned long long)s, acb_seed_epoch_ptr()}; }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// generic typed load/store through a runtime dtype tag (0 = f32, 1 = bf16)
__device__ __forceinline__ float ld_any(const void* p, long long i, int dt) {
  return dt == ACB_F32 ? ((const float*)p)[i] : __bfloat162float(((const bf16*)p)[i]);
}
__device__ __forceinline__ void st_any(void* p, long long i, int dt, float v) {
  if (dt == ACB_F32) ((float*)p)[i] = v;
  else ((bf16*)p)[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// erf-GELU through the Abramowitz-Stegun 7.1.26 rational (|erf err| <= 1.5e-7): ~12 instructions, used
// wherever the result is rounded to bf16 or feeds a bf16 tensor-core GEMM.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_z = poly * t * __expf(-z * z);  // 1 - erf(|x|/sqrt2)
  const float cdf = x >= 0.0f ? 1.0f - 0.5f * erfc_z : 0.5f * erfc_z;
  return x * cdf;
}
// bf16-output GELU: x * Phi(x) through the hardware tanh (MUFU.TANH), 7 instructions / 1 MUFU.  |error| <= 4.8e-4
// against the exact erf form, below the bf16 rounding of the O(1) activations it feeds; used only where the result is
// rounded to bf16 (define ACB_EXACT_GELU_BF16 to fall back to the 1.5e-7-accurate rational).
__device__ __forceinline__ float gelu_bf16(float x) {
#ifdef ACB_EXACT_GELU_BF16
  return gelu_fast(x);
#else
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);  // sqrt(2/pi) * (x + 0.044715 x^3)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
#endif
}
// two values at once on Blackwell's packed fp32 pipe (FFMA2 / FMUL2: one issue slot per PAIR of operations; the epilogues that call
// this are issue-bound).  Same operations in the same order as gelu_bf16: bit-identical results.
__device__ __forceinline__ float2 gelu_bf16x2(float2 x) {
#ifdef ACB_EXACT_GELU_BF16
  return make_float2(gelu_fast(x.x), gelu_fast(x.y));
#else
  const float2 u = __fmul2_rn(x, __ffma2_rn(make_float2(0.0356774081f, 0.0356774081f), __fmul2_rn(x, x), make_float2(0.7978845608f, 0.7978845608f)));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hx = __fmul2_rn(make_float2(0.5f, 0.5f), x);
  return __ffma2_rn(hx, t, hx);
#endif
}
// derivative of gelu_bf16 (the tanh form the bf16 forward evaluates): ~10 instructions against ~40 for the erf/exp form
__device__ __forceinline__ float gelu_bf16_grad(float x) {
#ifdef ACB_EXACT_GELU_BF16
  return gelu_erf_grad(x);
#else
  const float x2 = x * x;
  const float u = x * fmaf(0.0356774081f, x2, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);
  return fmaf(0.5f * x * (1.0f - t * t), du, 0.5f * (1.0f + t));
#endif
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACB_ACT_RELU: return fmaxf(v, 0.0f);
    case ACB_ACT_GELU: return gelu_erf(v);
    case ACB_ACT_TANH: return tanhf(v);
    case ACB_ACT_SIGMOID: return sigmoidf_(v);
    case ACB_ACT_SOFTPLUS: return v > 20.0f ? v : log1pf(expf(v));
    default: return v;
  }
}

// counter hash for the Time2Vec dropout of the masked pre-training path (same mask in forward and backward)
__device__ __forceinline__ unsigned te_hash(unsigned long long seed, int t, int c) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)(unsigned)t << 10) | (unsigned long long)c);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (result broadcast to all threads)
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  return sh[32];
}
