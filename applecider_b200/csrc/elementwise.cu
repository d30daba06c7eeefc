// HBM-bound row kernels: LayerNorm (+fusions), casts, weight packing, pooling, softmax.
// One warp per row, 128-bit (f32) / 64-bit (bf16) vector IO, warp-shuffle reductions.
#include "common.cuh"

namespace {

// ---- vector IO: 4 consecutive elements --------------------------------------------------------
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// ---- LayerNorm ---------------------------------------------------------------------------------
// Two-pass statistics (mean, then centred variance) from an L1-resident row; biased variance like torch.
template <typename TX, typename TR, typename TY, bool VEC>
__global__ void __launch_bounds__(256) layernorm_kernel(const TX* __restrict__ x, const TR* __restrict__ res,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        TY* __restrict__ y, long long rows, int C, float eps,
                                                        int pre_gelu, int post_act, const int* __restrict__ rows_dev) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (rows_dev) rows = min(rows, (long long)__ldg(rows_dev));  // device-side row count: capacity rows past it are skipped
  if (row >= rows) return;
  const TX* xr = x + row * C;
  const TR* rr = res ? res + row * C : nullptr;
  TY* yr = y + row * C;

  auto value = [&](int c) -> float {
    float v = to_f<TX>(xr[c]);
    if (pre_gelu) v = gelu_erf(v);
    if (rr) v += to_f<TR>(rr[c]);
    return v;
  };

  float s = 0.0f;
  if (VEC) {
    for (int c = lane * 4; c < C; c += 128) {
      float v[4];
      Vec4<TX>::load(xr + c, v);
      if (pre_gelu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
      }
      if (rr) {
        float r[4];
        Vec4<TR>::load(rr + c, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += r[i];
      }
      s += (v[0] + v[1]) + (v[2] + v[3]);
    }
  } else {
    for (int c = lane; c < C; c += 32) s += value(c);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
  if (VEC) {
    for (int c = lane * 4; c < C; c += 128) {
      float v[4];
      Vec4<TX>::load(xr + c, v);
      if (pre_gelu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
      }
      if (rr) {
        float r[4];
        Vec4<TR>::load(rr + c, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += r[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) q += (v[i] - mean) * (v[i] - mean);
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float d = value(c) - mean;
      q += d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  if (VEC) {
    for (int c = lane * 4; c < C; c += 128) {
      float v[4];
      Vec4<TX>::load(xr + c, v);
      if (pre_gelu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
      }
      if (rr) {
        float r[4];
        Vec4<TR>::load(rr + c, r);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += r[i];
      }
      const float4 w4 = *reinterpret_cast<const float4*>(w + c);
      const float4 b4 = *reinterpret_cast<const float4*>(b + c);
      float o[4];
      o[0] = (v[0] - mean) * rstd * w4.x + b4.x;
      o[1] = (v[1] - mean) * rstd * w4.y + b4.y;
      o[2] = (v[2] - mean) * rstd * w4.z + b4.z;
      o[3] = (v[3] - mean) * rstd * w4.w + b4.w;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = apply_act(o[i], post_act);
      Vec4<TY>::store(yr + c, o);
    }
  } else {
    for (int c = lane; c < C; c += 32) {
      const float o = (value(c) - mean) * rstd * w[c] + b[c];
      yr[c] = from_f<TY>(apply_act(o, post_act));
    }
  }
}

// Register-cached variant for C <= 128*NC (every row element is read from global exactly once);
// GELU uses the 1.5e-7-accurate rational when the result is rounded to bf16 anyway.
template <typename TX, typename TR, typename TY, int NC>
__global__ void __launch_bounds__(256) layernorm_cached_kernel(const TX* __restrict__ x, const TR* __restrict__ res,
                                                               const float* __restrict__ w, const float* __restrict__ b,
                                                               TY* __restrict__ y, long long rows, int C, float eps,
                                                               int pre_gelu, int post_act, const int* __restrict__ rows_dev) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (rows_dev) rows = min(rows, (long long)__ldg(rows_dev));
  if (row >= rows) return;
  const TX* xr = x + row * C;
  const TR* rr = res ? res + row * C : nullptr;
  TY* yr = y + row * C;
  constexpr bool FAST = sizeof(TY) == 2;
  float v[NC][4];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < C) {
      Vec4<TX>::load(xr + c, v[k]);
      if (pre_gelu) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[k][i] = FAST ? gelu_fast(v[k][i]) : gelu_erf(v[k][i]);
      }
      if (rr) {
        float r4[4];
        Vec4<TR>::load(rr + c, r4);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[k][i] += r4[i];
      }
      s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    if (lane * 4 + k * 128 < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) q += (v[k][i] - mean) * (v[k][i] - mean);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int c = lane * 4 + k * 128;
    if (c < C) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + c);
      const float4 b4 = *reinterpret_cast<const float4*>(b + c);
      float o[4];
      o[0] = (v[k][0] - mean) * rstd * w4.x + b4.x;
      o[1] = (v[k][1] - mean) * rstd * w4.y + b4.y;
      o[2] = (v[k][2] - mean) * rstd * w4.z + b4.z;
      o[3] = (v[k][3] - mean) * rstd * w4.w + b4.w;
      if (post_act == ACB_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = FAST ? gelu_bf16(o[i]) : gelu_erf(o[i]);
      } else if (post_act != ACB_ACT_NONE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = apply_act(o[i], post_act);
      }
      Vec4<TY>::store(yr + c, o);
    }
  }
}

// Streaming variant for the large bf16 activations (SpectraNet LayerNorm+GELU, >= 64 K rows): a warp owns RPW rows
// and issues the loads of ALL of them before touching the first (RPW x C x 2 bytes in flight per warp -- one row per
// warp leaves HBM latency-bound at ~2.6 TB/s), affine parameters live in registers across the rows.
// LPR = lanes per row: 32 (C = NC x 128), or 16 for C = NC x 64 rows such as the 192 channels of SpectraNet's first block -- a
// warp then works on two rows side by side and the reductions stay inside a half-warp.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NC, int RPW, int LPR = 32>
__global__ void __launch_bounds__(256) layernorm_stream_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ b, bf16* __restrict__ y, long long rows, float eps,
                                                               int post_act, const int* __restrict__ rows_dev) {
  constexpr int CW = LPR * 4, C = NC * CW, RS = 32 / LPR;  // chunk width, row length, rows side by side in a warp
  const int lane = threadIdx.x & (LPR - 1), sub = (threadIdx.x & 31) / LPR;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (RPW * RS) + sub;
  if (rows_dev) rows = min(rows, (long long)__ldg(rows_dev));
  if (row0 - sub >= rows) return;  // (warp-uniform: the shuffles below need every lane of a live warp)
  uint2 raw[RPW][NC];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long long row = row0 + r * RS < rows ? row0 + r * RS : rows - 1;  // clamp: tail rows recompute the last row, stores are guarded
#pragma unroll
    for (int k = 0; k < NC; ++k) raw[r][k] = __ldcs(reinterpret_cast<const uint2*>(x + row * C + lane * 4 + k * CW));
  }
  constexpr bool WREG = NC <= 6;  // wide rows re-read the affine parameters from L1 instead of pinning 8*NC registers
  float4 wv[WREG ? NC : 1], bv[WREG ? NC : 1];
  if constexpr (WREG) {
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      wv[k] = *reinterpret_cast<const float4*>(w + lane * 4 + k * CW);
      bv[k] = *reinterpret_cast<const float4*>(b + lane * 4 + k * CW);
    }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float v[NC][4];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw[r][k].x);
      const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&raw[r][k].y);
      v[k][0] = __low2float(a); v[k][1] = __high2float(a); v[k][2] = __low2float(c); v[k][3] = __high2float(c);
      s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
    }
    const float mean = group_sum<LPR>(s) / (float)C;  // LPR = 32: same arithmetic as layernorm_cached_kernel, results do not depend on the row count
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q += (v[k][i] - mean) * (v[k][i] - mean);
      }
    const float rstd = rsqrtf(group_sum<LPR>(q) / (float)C + eps);
    if (row0 + r * RS < rows) {
      bf16* yr = y + (row0 + r * RS) * C;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        float o[4];
        const float4 w4 = WREG ? wv[WREG ? k : 0] : *reinterpret_cast<const float4*>(w + lane * 4 + k * CW);
        const float4 b4 = WREG ? bv[WREG ? k : 0] : *reinterpret_cast<const float4*>(b + lane * 4 + k * CW);
        o[0] = (v[k][0] - mean) * rstd * w4.x + b4.x;
        o[1] = (v[k][1] - mean) * rstd * w4.y + b4.y;
        o[2] = (v[k][2] - mean) * rstd * w4.z + b4.z;
        o[3] = (v[k][3] - mean) * rstd * w4.w + b4.w;
        if (post_act == ACB_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = gelu_bf16(o[i]);
        } else if (post_act != ACB_ACT_NONE) {
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = apply_act(o[i], post_act);
        }
        __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t*>(&h0);
        t.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(yr + lane * 4 + k * CW) = t;
      }
    }
  }
}

template <typename TX, typename TR, typename TY>
int launch_ln(const void* x, const void* res, const float* w, const float* b, void* y, long long rows, int C,
              float eps, int pre_gelu, int post_act, const int* rows_dev, cudaStream_t st) {
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  const bool vec = (C % 4 == 0) && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)res | (uintptr_t)w | (uintptr_t)b) % 16 == 0);
  if constexpr (sizeof(TX) == 2 && sizeof(TY) == 2) {
    if (vec && !res && !pre_gelu && rows >= 65536 && (C == 128 || C == 192 || C == 256 || C == 384 || C == 768 || C == 1536 || C == 3072)) {
      constexpr int RPW = 4;
      const unsigned g = (unsigned)((rows + (long long)wpb * RPW - 1) / ((long long)wpb * RPW));
      if (C == 192) layernorm_stream_kernel<3, 4, 16><<<(unsigned)((rows + wpb * 8LL - 1) / (wpb * 8LL)), wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      else if (C == 128) layernorm_stream_kernel<1, 8><<<(unsigned)((rows + wpb * 8LL - 1) / (wpb * 8LL)), wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      else if (C == 256) layernorm_stream_kernel<2, RPW><<<g, wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      else if (C == 384) layernorm_stream_kernel<3, RPW><<<g, wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      else if (C == 768) layernorm_stream_kernel<6, 2><<<(unsigned)((rows + wpb * 2LL - 1) / (wpb * 2LL)), wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      else if (C == 1536) layernorm_stream_kernel<12, 1><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      else layernorm_stream_kernel<24, 1><<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>((const bf16*)x, w, b, (bf16*)y, rows, eps, post_act, rows_dev);
      ACB_LAUNCH_CHECK();
      acb_count_launch();
      return ACB_OK;
    }
  }
  if (vec && C <= 256)
    layernorm_cached_kernel<TX, TR, TY, 2><<<grid, wpb * 32, 0, st>>>((const TX*)x, (const TR*)res, w, b, (TY*)y, rows, C, eps, pre_gelu, post_act, rows_dev);
  else if (vec && C <= 768)
    layernorm_cached_kernel<TX, TR, TY, 6><<<grid, wpb * 32, 0, st>>>((const TX*)x, (const TR*)res, w, b, (TY*)y, rows, C, eps, pre_gelu, post_act, rows_dev);
  else if (vec)
    layernorm_kernel<TX, TR, TY, true><<<grid, wpb * 32, 0, st>>>((const TX*)x, (const TR*)res, w, b, (TY*)y, rows, C, eps, pre_gelu, post_act, rows_dev);
  else
    layernorm_kernel<TX, TR, TY, false><<<grid, wpb * 32, 0, st>>>((const TX*)x, (const TR*)res, w, b, (TY*)y, rows, C, eps, pre_gelu, post_act, rows_dev);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

// ---- casts / packing ---------------------------------------------------------------------------
__global__ void cast_kernel(const void* in, int idt, void* out, int odt, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    st_any(out, i, odt, ld_any(in, i, idt));
}

__global__ void pack_conv_weight_kernel(const float* w, void* out, int odt, int Cout, int Cin, int k, long long row_stride,
                                        int tap_off) {
  const long long total = (long long)Cout * Cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // iterate in OUTPUT order (co, tap, ci) for coalesced writes
    const int ci = (int)(i % Cin);
    const long long t = i / Cin;
    const int tap = (int)(t % k);
    const int co = (int)(t / k);
    const float v = __ldg(w + ((long long)co * Cin + ci) * k + tap);
    st_any(out, (long long)co * row_stride + (long long)(tap + tap_off) * Cin + ci, odt, v);
  }
}

__global__ void pack_polyphase_kernel(const float* w, void* out, int odt, int Cout, int k, int phases, int halo,
                                      int rows_per_phase, int row_off, long long row_stride, int Kp) {
  const long long total = (long long)phases * Cout * Kp;
  const int pad = k / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % Kp);
    const long long t = i / Kp;
    const int co = (int)(t % Cout);
    const int r = (int)(t / Cout);
    const int tap = kk - r + pad - halo;
    const float v = (tap >= 0 && tap < k) ? __ldg(w + (long long)co * k + tap) : 0.0f;
    st_any(out, ((long long)row_off + (long long)r * rows_per_phase + co) * row_stride + kk, odt, v);
  }
}

__global__ void pack_conv2d_weight_kernel(const float* w, void* out, int odt, int Cout, int Cin, int kh, int kw) {
  const long long total = (long long)Cout * Cin * kh * kw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    long long t = i / Cin;
    const int kx = (int)(t % kw);
    t /= kw;
    const int ky = (int)(t % kh);
    const int co = (int)(t / kh);
    st_any(out, i, odt, __ldg(w + (((long long)co * Cin + ci) * kh + ky) * kw + kx));
  }
}

__global__ void pad_signal_kernel(const float* in, void* out, int odt, int nb, int L, long long out_stride, int lead) {
  const long long total = (long long)nb * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / L;
    const int l = (int)(i - b * L);
    st_any(out, b * out_stride + lead + l, odt, __ldg(in + i));
  }
}

// out[c*R + r] = in[r*C + c]  (32x32 smem tiles, dtype-tagged; used for bf16 W^T copies of the dgrad GEMMs)
__global__ void __launch_bounds__(256) transpose_kernel(const void* in, int idt, void* out, int odt, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < R && c < C) tile[i][tx] = ld_any(in, (long long)r * C + c, idt);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < C) st_any(out, (long long)c * R + r, odt, tile[tx][i]);
  }
}

// The same for MANY fp32 matrices in one launch (all the W^T bf16 copies the dgrad GEMMs of a training step need, refreshed
// together after the optimizer step): job j = (src, dst, R, C) owns the global tiles [tile0[j], tile0[j+1]), a block finds its job
// by binary search.  meta[j] = {R, C, tiles per row, first tile}.
__global__ void __launch_bounds__(256) transpose_batch_kernel(const long long* __restrict__ src, const long long* __restrict__ dst,
                                                              const int4* __restrict__ meta, int n_jobs) {
  __shared__ float tile[32][33];
  const int t = blockIdx.x;
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {  // last job whose first tile is <= t
    const int mid = (lo + hi + 1) >> 1;
    if (meta[mid].w <= t) lo = mid; else hi = mid - 1;
  }
  const int4 m = meta[lo];
  const int R = m.x, C = m.y, lt = t - m.w;
  const int c0 = (lt % m.z) * 32, r0 = (lt / m.z) * 32;
  const float* in = reinterpret_cast<const float*>(src[lo]);
  bf16* out = reinterpret_cast<bf16*>(dst[lo]);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < R && c < C) tile[i][tx] = in[(long long)r * C + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < C) out[(long long)c * R + r] = __float2bfloat16(tile[tx][i]);
  }
}

// ---- pooling over L of channels-last [B, L, C] ---------------------------------------------------
template <typename T>
__global__ void maxpool4_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int L, int C) {
  const int Lo = L / 4;
  const long long total = (long long)B * Lo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int lo = (int)(t % Lo);
    const long long b = t / Lo;
    const T* p = x + ((b * L + 4LL * lo) * C + c);
    float m = to_f<T>(p[0]);
    m = fmaxf(m, to_f<T>(p[C]));
    m = fmaxf(m, to_f<T>(p[2LL * C]));
    m = fmaxf(m, to_f<T>(p[3LL * C]));
    y[i] = from_f<T>(m);
  }
}

template <typename T>
__global__ void globalmax_kernel(const T* __restrict__ x, float* __restrict__ y, int B, int L, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const int c = (int)(i % C);
  const long long b = i / C;
  const T* p = x + b * L * C + c;
  float m = -INFINITY;
  for (int l = 0; l < L; ++l) m = fmaxf(m, to_f<T>(p[(long long)l * C]));
  y[i] = m;
}

__global__ void softmax_rows_kernel(const float* x, float* y, int rows, int C) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, x[(long long)r * C + c]);
  float s = 0.0f;
  for (int c = 0; c < C; ++c) s += expf(x[(long long)r * C + c] - m);
  for (int c = 0; c < C; ++c) y[(long long)r * C + c] = expf(x[(long long)r * C + c] - m) / s;
}

inline unsigned grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int acb_layernorm(const void* x, int x_dtype, const void* res, int res_dtype, const float* w, const float* b, void* y,
                  int y_dtype, long long rows, int C, float eps, int pre_gelu, int post_act, void* stream) {
  return acb_layernorm_n(x, x_dtype, res, res_dtype, w, b, y, y_dtype, rows, C, eps, pre_gelu, post_act, nullptr, stream);
}

int acb_layernorm_n(const void* x, int x_dtype, const void* res, int res_dtype, const float* w, const float* b, void* y,
                    int y_dtype, long long rows, int C, float eps, int pre_gelu, int post_act, const int* rows_dev, void* stream) {
  ACB_CHECK(x && y && w && b && C > 0 && rows >= 0, "acb_layernorm: bad arguments");
  if (rows == 0) return ACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!res) res_dtype = x_dtype;
  const int key = x_dtype * 4 + res_dtype * 2 + y_dtype;
  switch (key) {
    case 0: return launch_ln<float, float, float>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 1: return launch_ln<float, float, bf16>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 2: return launch_ln<float, bf16, float>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 3: return launch_ln<float, bf16, bf16>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 4: return launch_ln<bf16, float, float>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 5: return launch_ln<bf16, float, bf16>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 6: return launch_ln<bf16, bf16, float>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
    case 7: return launch_ln<bf16, bf16, bf16>(x, res, w, b, y, rows, C, eps, pre_gelu, post_act, rows_dev, st);
  }
  acb_set_error("acb_layernorm: bad dtype");
  return ACB_ERR_INVALID;
}

int acb_cast(const void* in, int in_dtype, void* out, int out_dtype, long long n, void* stream) {
  ACB_CHECK(in && out && n >= 0, "acb_cast: bad arguments");
  if (n == 0) return ACB_OK;
  cast_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, in_dtype, out, out_dtype, n);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_transpose(const void* in, int in_dtype, void* out, int out_dtype, int R, int C, void* stream) {
  ACB_CHECK(in && out && R > 0 && C > 0, "acb_transpose: bad arguments");
  dim3 grid(cdiv(C, 32), cdiv(R, 32));
  ACB_CHECK(grid.y <= 65535, "acb_transpose: too many rows");
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, in_dtype, out, out_dtype, R, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_transpose_batch(const long long* src_ptrs, const long long* dst_ptrs, const int* meta, int n_jobs, int total_tiles, void* stream) {
  ACB_CHECK(src_ptrs && dst_ptrs && meta && n_jobs > 0 && total_tiles > 0, "acb_transpose_batch: bad arguments");
  transpose_batch_kernel<<<total_tiles, 256, 0, (cudaStream_t)stream>>>(src_ptrs, dst_ptrs, reinterpret_cast<const int4*>(meta), n_jobs);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_pack_conv_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int k, long long row_stride,
                         int tap_off, void* stream) {
  ACB_CHECK(w && out && Cout > 0 && Cin > 0 && k > 0, "acb_pack_conv_weight: bad arguments");
  pack_conv_weight_kernel<<<grid_for((long long)Cout * Cin * k), 256, 0, (cudaStream_t)stream>>>(w, out, out_dtype, Cout, Cin, k, row_stride, tap_off);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_pack_polyphase_weight(const float* w, void* out, int out_dtype, int Cout, int k, int phases, int halo,
                              int rows_per_phase, int row_off, long long row_stride, int Kp, void* stream) {
  ACB_CHECK(w && out && Cout > 0 && k > 0 && phases > 0 && Kp > 0, "acb_pack_polyphase_weight: bad arguments");
  pack_polyphase_kernel<<<grid_for((long long)phases * Cout * Kp), 256, 0, (cudaStream_t)stream>>>(w, out, out_dtype, Cout, k, phases, halo, rows_per_phase, row_off, row_stride, Kp);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_pack_conv2d_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int kh, int kw, void* stream) {
  ACB_CHECK(w && out, "acb_pack_conv2d_weight: bad arguments");
  pack_conv2d_weight_kernel<<<grid_for((long long)Cout * Cin * kh * kw), 256, 0, (cudaStream_t)stream>>>(w, out, out_dtype, Cout, Cin, kh, kw);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_pad_signal(const float* in, void* out, int out_dtype, int nb, int L, long long out_stride, int lead, void* stream) {
  ACB_CHECK(in && out && nb > 0 && L > 0, "acb_pad_signal: bad arguments");
  pad_signal_kernel<<<grid_for((long long)nb * L), 256, 0, (cudaStream_t)stream>>>(in, out, out_dtype, nb, L, out_stride, lead);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_maxpool4_cl(const void* x, int dtype, void* y, int B, int L, int C, void* stream) {
  ACB_CHECK(x && y && B > 0 && L >= 4 && C > 0, "acb_maxpool4_cl: bad arguments");
  const long long n = (long long)B * (L / 4) * C;
  if (dtype == ACB_F32)
    maxpool4_kernel<float><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, B, L, C);
  else
    maxpool4_kernel<bf16><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, L, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

// pair max of bf16 rows, 16 bytes per thread: y[r, :] = max(x[2r, :], x[2r+1, :])  (second half of MaxPool1d(4) after the
// stage-0 kernel has already max-ed the two phases it owns)
__global__ void pairmax_bf16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long rows_out, int vec_per_row) {
  const long long total = rows_out * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row);
    const uint4 a = __ldcs(x + (2 * r) * vec_per_row + c), b = __ldcs(x + (2 * r + 1) * vec_per_row + c);
    uint4 o;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) po[k] = __hmax2(pa[k], pb[k]);
    y[i] = o;
  }
}

// legacy spectra encoder (brew_cider.py:611-636): cat([MaxPool1d(4), AvgPool1d(4), -MaxPool1d(4)(-x)]) on channels-last fp32
__global__ void tripool4_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int L, int C) {
  const int Lo = L / 4;
  const long long total = (long long)B * Lo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int lo = (int)(t % Lo);
    const long long b = t / Lo;
    const float* p = x + ((b * L + 4LL * lo) * C + c);
    const float v0 = p[0], v1 = p[C], v2 = p[2LL * C], v3 = p[3LL * C];
    float* o = y + (b * Lo + lo) * (3LL * C) + c;
    o[0] = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    o[C] = (((v0 + v1) + v2) + v3) * 0.25f;
    o[2LL * C] = fminf(fminf(v0, v1), fminf(v2, v3));
  }
}

int acb_tripool4_cl(const float* x, float* y, int B, int L, int C, void* stream) {
  ACB_CHECK(x && y && B > 0 && L >= 4 && C > 0, "acb_tripool4_cl: bad arguments");
  tripool4_kernel<<<grid_for((long long)B * (L / 4) * C), 256, 0, (cudaStream_t)stream>>>(x, y, B, L, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_pairmax_bf16(const void* x, void* y, long long rows_out, int C, void* stream) {
  ACB_CHECK(x && y && rows_out >= 0 && C > 0 && C % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), "acb_pairmax_bf16: bad arguments");
  if (rows_out == 0) return ACB_OK;
  pairmax_bf16_kernel<<<grid_for(rows_out * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, rows_out, C / 8);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_globalmax_cl(const void* x, int dtype, float* y, int B, int L, int C, void* stream) {
  ACB_CHECK(x && y && B > 0 && L > 0 && C > 0, "acb_globalmax_cl: bad arguments");
  const long long n = (long long)B * C;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == ACB_F32)
    globalmax_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, y, B, L, C);
  else
    globalmax_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, y, B, L, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_softmax_rows(const float* x, float* y, int rows, int C, void* stream) {
  ACB_CHECK(x && y && rows > 0 && C > 0, "acb_softmax_rows: bad arguments");
  softmax_rows_kernel<<<cdiv(rows, 128), 128, 0, (cudaStream_t)stream>>>(x, y, rows, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
