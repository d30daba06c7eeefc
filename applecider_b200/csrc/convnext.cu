// ConvNeXt-T support kernels (channels-last): stem patchify, depthwise 7x7 + LayerNorm,
// LayerNorm2d + 2x2 patch gather for the downsample convs, global-average-pool + LayerNorm head.
// The dense parts (stem / MLP / downsample) are GEMMs (gemm_f32.cu / gemm_tc.cu).
#include "common.cuh"

namespace {

template <typename T>
__global__ void patchify_kernel(const float* __restrict__ img, int B, int Cin, int H, int W, int p, T* __restrict__ out) {
  const int Ho = H / p, Wo = W / p, K = Cin * p * p;
  const long long total = (long long)B * Ho * Wo * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    long long m = i / K;
    const int ox = (int)(m % Wo);
    m /= Wo;
    const int oy = (int)(m % Ho);
    const int b = (int)(m / Ho);
    const int kx = k % p, ky = (k / p) % p, ci = k / (p * p);
    out[i] = from_f<T>(__ldg(img + (((long long)b * Cin + ci) * H + (oy * p + ky)) * W + (ox * p + kx)));
  }
}

// p = 4, bf16 output (the ConvNeXt stem): one thread per (pixel, channel, ky) moves 4 kx values -- consecutive threads walk
// (ci, ky) so a warp writes contiguous 8-byte pieces of the patch rows (the per-element kernel above is 6x off the HBM bound)
__global__ void __launch_bounds__(256) patchify4_bf16_kernel(const float* __restrict__ img, int B, int Cin, int H, int W, bf16* __restrict__ out) {
  const int Ho = H / 4, Wo = W / 4, G = Cin * 4;  // (ci, ky) groups per pixel
  const long long total = (long long)B * Ho * Wo * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long m = i / G;
    const int ox = (int)(m % Wo);
    const long long t = m / Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int ci = g >> 2, ky = g & 3;
    const float* src = img + (((long long)b * Cin + ci) * H + (oy * 4 + ky)) * W + ox * 4;
    const float v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out + m * (Cin * 16) + g * 4) = o;
  }
}

// One CTA per (image, output row). Input rows oy-3..oy+3 are staged in shared memory (fp32, zero
// padded), every thread produces (ox, c) outputs, then one warp per pixel applies LayerNorm over C.
template <typename T>
__global__ void __launch_bounds__(256) dwconv7_ln_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                         const float* __restrict__ ln_b, float eps, T* __restrict__ y,
                                                         int H, int W, int C) {
  extern __shared__ float sm[];
  float* in = sm;                         // [7][W][C]
  float* cv = sm + (size_t)7 * W * C;     // [W][C]
  const int b = blockIdx.x / H, oy = blockIdx.x % H;
  const int WC = W * C;
  for (int i = threadIdx.x; i < 7 * WC; i += blockDim.x) {
    const int r = i / WC, rem = i - r * WC;
    const int iy = oy + r - 3;
    in[i] = (iy >= 0 && iy < H) ? to_f<T>(x[((long long)b * H + iy) * WC + rem]) : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < WC; i += blockDim.x) {
    const int ox = i / C, c = i - ox * C;
    const float* wc = w + c * 49;
    float acc = bias[c];
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int ix = ox + kx - 3;
        if (ix >= 0 && ix < W) acc = fmaf(in[(ky * W + ix) * C + c], __ldg(wc + ky * 7 + kx), acc);
      }
    }
    cv[i] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int ox = wid; ox < W; ox += nw) {
    const float* v = cv + ox * C;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += v[c];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float d = v[c] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    T* o = y + (((long long)b * H + oy) * W + ox) * C;
    for (int c = lane; c < C; c += 32) o[c] = from_f<T>((v[c] - mean) * rstd * ln_w[c] + ln_b[c]);
  }
}

// v2 (W known at compile time): thread = channel, CTA = (image, strip of R output rows).  The 49 taps live in
// registers, one input row segment (W values) is loaded per ky and reused by all W x 7 taps (register
// sliding window), so shared-memory traffic drops 5x against the generic kernel; out-of-range taps vanish
// at compile time.  LayerNorm: conv row -> smem -> warp per pixel, channel pairs per lane.
// LayerNorm of the W pixels of one conv row held in shared memory (cv[W][C] fp32): warp per pixel, NPL = C/32 values
// per lane in registers (read once), affine parameters from shared memory, two-pass statistics.
// A warp works on PB pixels AT ONCE: the two butterfly reductions of a pixel are 10 dependent shuffles, and with one pixel in
// flight per warp (3 warps per CTA at C = 96) the LayerNorm phase was 45 % of the kernel's stall samples (ncu, short scoreboard).
template <typename T, int NPL, int PB>
__device__ __forceinline__ void dw_ln_row(const float* __restrict__ cv, const float* __restrict__ s_lnw, const float* __restrict__ s_lnb,
                                          float eps, T* __restrict__ yrow, int Wpix, int wid, int nw, int lane) {
  constexpr int C = NPL * 32;
  for (int ox0 = wid; ox0 < Wpix; ox0 += nw * PB) {
    float t[PB][NPL], s[PB], q[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      const int ox = min(ox0 + p * nw, Wpix - 1);  // (clamped pixels are recomputed, their stores are skipped)
      const float* v = cv + ox * C;
      s[p] = 0.0f;
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        t[p][k] = v[lane + 32 * k];
        s[p] += t[p][k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int p = 0; p < PB; ++p) s[p] += __shfl_xor_sync(0xffffffffu, s[p], o);
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      const float mean = s[p] * (1.0f / (float)C);
      q[p] = 0.0f;
#pragma unroll
      for (int k = 0; k < NPL; ++k) {
        t[p][k] -= mean;
        q[p] = fmaf(t[p][k], t[p][k], q[p]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int p = 0; p < PB; ++p) q[p] += __shfl_xor_sync(0xffffffffu, q[p], o);
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      const int ox = ox0 + p * nw;
      if (ox < Wpix) {
        const float rstd = rsqrtf(q[p] * (1.0f / (float)C) + eps);
        T* o = yrow + (long long)ox * C;
#pragma unroll
        for (int k = 0; k < NPL; ++k) o[lane + 32 * k] = from_f<T>(fmaf(t[p][k] * rstd, s_lnw[lane + 32 * k], s_lnb[lane + 32 * k]));
      }
    }
  }
}

template <int W>
constexpr int dw_max_threads() { return W >= 15 ? 256 : (W >= 7 ? 512 : 768); }

template <typename T, int W>
__global__ void __launch_bounds__(dw_max_threads<W>()) dwconv7_ln_w_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                           const float* __restrict__ ln_b, float eps, T* __restrict__ y,
                                                           int B, int H, int C, int R, int ipb, int use_wsm, int do_ln, int flip) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int strips = (H + R - 1) / R;
  const int b0 = (blockIdx.x / strips) * ipb, oy0 = (blockIdx.x % strips) * R;  // ipb > 1 only when one strip covers the image
  const int rows = min(R, H - oy0);
  const int WC = W * C;
  T* in = reinterpret_cast<T*>(smraw);
  const size_t in_bytes = (((size_t)(R + 6) * WC * sizeof(T)) + 15) & ~(size_t)15;
  constexpr bool STRIP_LN = W <= 3;  // small maps: LayerNorm once per strip over all its R x W pixels (one pixel per warp and row leaves most warps idle)
  float* cv = reinterpret_cast<float*>(smraw + in_bytes);
  float* wsm = cv + (STRIP_LN ? R * WC : WC);
  float* s_lnw = wsm + (use_wsm ? 49 * (C + 1) : 0);
  float* s_lnb = s_lnw + C;
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (do_ln)
    for (int i = tid; i < C; i += nthr) {
      s_lnw[i] = ln_w[i];
      s_lnb[i] = ln_b[i];
    }
  if (use_wsm) {  // straight coalesced copy [C][49]: the odd row stride makes the per-thread reads below conflict-free
    for (int i = tid; i < 49 * C; i += nthr) wsm[i] = __ldg(w + i);
    __syncthreads();
  }
  const int c = tid;
  float wr[49];
  float bc = 0.0f;
  if (c < C) {  // the 49 taps of this thread's channel stay in registers for every image of the CTA
#pragma unroll
    for (int j = 0; j < 49; ++j) wr[j] = use_wsm ? wsm[c * 49 + (flip ? 48 - j : j)] : __ldg(w + c * 49 + (flip ? 48 - j : j));
    bc = bias ? bias[c] : 0.0f;
  }
  const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  for (int bi = 0; bi < ipb; ++bi) {
  const int b = b0 + bi;
  if (b >= B) break;  // uniform
  {  // stage input rows oy0-3 .. oy0+rows+2 (16-byte copies; rows outside the image are zero)
    const int v_per_row = (int)((size_t)WC * sizeof(T) / 16);
    const uint4* xg = reinterpret_cast<const uint4*>(x + (long long)b * H * WC);
    uint4* ins = reinterpret_cast<uint4*>(in);
    const int nvec = (rows + 6) * v_per_row;
    for (int i = tid; i < nvec; i += nthr) {  // flat loop: many independent loads in flight per thread
      const int rr = i / v_per_row, rem = i - rr * v_per_row;
      const int iy = oy0 + rr - 3;
      ins[i] = (iy >= 0 && iy < H) ? __ldg(xg + (long long)iy * v_per_row + rem) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  for (int r = 0; r < rows; ++r) {
    if (c < C) {
      float acc[W];
#pragma unroll
      for (int ox = 0; ox < W; ++ox) acc[ox] = bc;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        float iv[W];
        const T* rowp = in + (size_t)(r + ky) * WC + c;
#pragma unroll
        for (int ix = 0; ix < W; ++ix) iv[ix] = to_f<T>(rowp[ix * C]);
#pragma unroll
        for (int ox = 0; ox < W; ++ox) {
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) {
            const int ix = ox + kx - 3;
            if (ix >= 0 && ix < W) acc[ox] = fmaf(iv[ix], wr[ky * 7 + kx], acc[ox]);
          }
        }
      }
      if (!do_ln) {
        T* o = y + (((long long)b * H + oy0 + r) * W) * C + c;
#pragma unroll
        for (int ox = 0; ox < W; ++ox) o[ox * C] = from_f<T>(acc[ox]);
      } else {
#pragma unroll
        for (int ox = 0; ox < W; ++ox) cv[(STRIP_LN ? r * WC : 0) + ox * C + c] = acc[ox];
      }
    }
    if (!do_ln) continue;  // uniform for the whole CTA
    if (STRIP_LN && (C == 384 || C == 768)) {
      if (r + 1 < rows) continue;
      __syncthreads();
      T* ybase = y + (((long long)b * H + oy0) * W) * C;  // the strip's rows are contiguous in y
      if (C == 384) dw_ln_row<T, 12, 1>(cv, s_lnw, s_lnb, eps, ybase, rows * W, wid, nw, lane);
      else dw_ln_row<T, 24, 1>(cv, s_lnw, s_lnb, eps, ybase, rows * W, wid, nw, lane);
      __syncthreads();
      continue;
    }
    __syncthreads();
    {
      T* yrow = y + (((long long)b * H + oy0 + r) * W) * C;
      bool done = true;
      switch (C) {
        case 96: dw_ln_row<T, 3, 5>(cv + (STRIP_LN ? r * WC : 0), s_lnw, s_lnb, eps, yrow, W, wid, nw, lane); break;
        case 192: dw_ln_row<T, 6, 2>(cv + (STRIP_LN ? r * WC : 0), s_lnw, s_lnb, eps, yrow, W, wid, nw, lane); break;
        case 384: dw_ln_row<T, 12, 1>(cv + (STRIP_LN ? r * WC : 0), s_lnw, s_lnb, eps, yrow, W, wid, nw, lane); break;
        case 768: dw_ln_row<T, 24, 1>(cv + (STRIP_LN ? r * WC : 0), s_lnw, s_lnb, eps, yrow, W, wid, nw, lane); break;
        default: done = false;
      }
      if (done) {
        __syncthreads();
        continue;
      }
    }
    for (int ox = wid; ox < W; ox += nw) {
      const float* v = cv + (STRIP_LN ? r * WC : 0) + ox * C;
      float s = 0.0f;
      for (int cc = lane * 2; cc < C; cc += 64) {
        const float2 t = *reinterpret_cast<const float2*>(v + cc);
        s += t.x + t.y;
      }
      const float mean = warp_sum(s) / (float)C;
      float q = 0.0f;
      for (int cc = lane * 2; cc < C; cc += 64) {
        const float2 t = *reinterpret_cast<const float2*>(v + cc);
        q += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
      T* o = y + (((long long)b * H + oy0 + r) * W + ox) * C;
      for (int cc = lane * 2; cc < C; cc += 64) {
        const float2 t = *reinterpret_cast<const float2*>(v + cc);
        const float2 g = *reinterpret_cast<const float2*>(ln_w + cc);
        const float2 be = *reinterpret_cast<const float2*>(ln_b + cc);
        const float o0 = (t.x - mean) * rstd * g.x + be.x, o1 = (t.y - mean) * rstd * g.y + be.y;
        if (sizeof(T) == 2) {
          *reinterpret_cast<__nv_bfloat162*>(o + cc) = __floats2bfloat162_rn(o0, o1);
        } else {
          *reinterpret_cast<float2*>(o + cc) = make_float2(o0, o1);
        }
      }
    }
    __syncthreads();
  }
  if (!do_ln) __syncthreads();  // the next image overwrites the staged rows
  }
}

template <typename T, int W>
int launch_dwconv_w(const void* x, const float* w, const float* b, const float* ln_w, const float* ln_b, float eps, void* y,
                    int B, int H, int C, cudaStream_t st, int do_ln = 1, int flip = 0) {
  const int R = W >= 15 ? 5 : (W >= 7 ? 7 : (W >= 3 ? 3 : 1));
  const int use_wsm = 0;  // measured: reading the 49 taps straight from L2 beats staging them in smem (19 KB less smem -> 5 CTAs/SM: 0.54 -> 0.46 ms)
  const size_t in_bytes = (((size_t)(R + 6) * W * C * sizeof(T)) + 15) & ~(size_t)15;
  const size_t smem = in_bytes + (size_t)(W <= 3 ? R : 1) * W * C * 4 + (use_wsm ? (size_t)49 * (C + 1) * 4 : 0) + (size_t)2 * C * 4;
  auto k = dwconv7_ln_w_kernel<T, W>;
  ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int threads = ((C + 31) / 32) * 32;
  if (threads > dw_max_threads<W>() || smem > 200 * 1024) return 1;  // fall back to the generic kernel
  const int strips = (H + R - 1) / R;
  // small maps: several images per CTA so the 49 taps per channel are fetched once per CTA, not once per image
  int ipb = 1;
  if (strips == 1) while (ipb < 16 && (long long)(B / (ipb * 2)) >= 148 * 4) ipb *= 2;
  const unsigned grid = (unsigned)((long long)((B + ipb - 1) / ipb) * strips);
  k<<<grid, threads, smem, st>>>((const T*)x, w, b, ln_w, ln_b, eps, (T*)y, B, H, C, R, ipb, use_wsm, do_ln, flip);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

template <typename T>
int dispatch_dwconv_w(int W, const void* x, const float* w, const float* b, const float* ln_w, const float* ln_b, float eps,
                      void* y, int B, int H, int C, cudaStream_t st, int do_ln = 1, int flip = 0) {
  switch (W) {
    case 15: return launch_dwconv_w<T, 15>(x, w, b, ln_w, ln_b, eps, y, B, H, C, st, do_ln, flip);
    case 7: return launch_dwconv_w<T, 7>(x, w, b, ln_w, ln_b, eps, y, B, H, C, st, do_ln, flip);
    case 3: return launch_dwconv_w<T, 3>(x, w, b, ln_w, ln_b, eps, y, B, H, C, st, do_ln, flip);
    case 1: return launch_dwconv_w<T, 1>(x, w, b, ln_w, ln_b, eps, y, B, H, C, st, do_ln, flip);
  }
  return 1;  // not specialised
}

// warp per INPUT pixel inside the floor(H/2) x floor(W/2) region
template <typename T>
__global__ void __launch_bounds__(256) ln_patch2_kernel(const T* __restrict__ x, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, float eps, T* __restrict__ out,
                                                        int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)B * Ho * 2 * Wo * 2;
  if (pix >= total) return;
  const int ix = (int)(pix % (Wo * 2));
  long long t = pix / (Wo * 2);
  const int iy = (int)(t % (Ho * 2));
  const int b = (int)(t / (Ho * 2));
  const T* v = x + (((long long)b * H + iy) * W + ix) * C;
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s += to_f<T>(v[c]);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float d = to_f<T>(v[c]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  const int oy = iy >> 1, ox = ix >> 1, ky = iy & 1, kx = ix & 1;
  T* o = out + (((long long)b * Ho + oy) * Wo + ox) * (4LL * C) + (ky * 2 + kx) * C;
  for (int c = lane; c < C; c += 32) o[c] = from_f<T>((to_f<T>(v[c]) - mean) * rstd * ln_w[c] + ln_b[c]);
}

// bf16 fast path: a warp owns PPW consecutive pixels and loads all of them (C/32 values per lane each) before the first
// reduction -- one pixel per warp left this kernel latency / CTA-launch bound at 7x the HBM time.
template <int NPL, int PPW>
__global__ void __launch_bounds__(256) ln_patch2_stream_kernel(const bf16* __restrict__ x, const float* __restrict__ ln_w,
                                                               const float* __restrict__ ln_b, float eps, bf16* __restrict__ out,
                                                               int B, int H, int W) {
  constexpr int C = NPL * 32;
  const int Ho = H / 2, Wo = W / 2;
  const int lane = threadIdx.x & 31;
  const long long pix0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PPW;
  const long long total = (long long)B * Ho * 2 * Wo * 2;
  if (pix0 >= total) return;
  float v[PPW][NPL];
  long long obase[PPW];
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp) {
    const long long pix = pix0 + pp < total ? pix0 + pp : total - 1;
    const int ix = (int)(pix % (Wo * 2));
    const long long t = pix / (Wo * 2);
    const int iy = (int)(t % (Ho * 2));
    const int b = (int)(t / (Ho * 2));
    const bf16* src = x + (((long long)b * H + iy) * W + ix) * C;
#pragma unroll
    for (int k = 0; k < NPL; ++k) v[pp][k] = __bfloat162float(src[lane + 32 * k]);
    obase[pp] = (((long long)b * Ho + (iy >> 1)) * Wo + (ix >> 1)) * (4LL * C) + ((iy & 1) * 2 + (ix & 1)) * C;
  }
  float wv[NPL], bv[NPL];
#pragma unroll
  for (int k = 0; k < NPL; ++k) {
    wv[k] = ln_w[lane + 32 * k];
    bv[k] = ln_b[lane + 32 * k];
  }
#pragma unroll
  for (int pp = 0; pp < PPW; ++pp) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) s += v[pp][k];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
      const float d = v[pp][k] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (pix0 + pp < total) {
      bf16* o = out + obase[pp];
#pragma unroll
      for (int k = 0; k < NPL; ++k) o[lane + 32 * k] = __float2bfloat16_rn((v[pp][k] - mean) * rstd * wv[k] + bv[k]);
    }
  }
}

// CTA per image: mean over HW per channel, then LayerNorm over C
template <typename T>
__global__ void __launch_bounds__(256) gap_ln_kernel(const T* __restrict__ x, const float* __restrict__ ln_w,
                                                     const float* __restrict__ ln_b, float eps, float* __restrict__ out,
                                                     int HW, int C) {
  extern __shared__ float sm[];  // [C] + 33
  float* mean_c = sm;
  float* red = sm + C;
  const int b = blockIdx.x;
  const T* xb = x + (long long)b * HW * C;
  float part = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f;
    for (int i = 0; i < HW; ++i) s += to_f<T>(xb[(long long)i * C + c]);
    s /= (float)HW;
    mean_c[c] = s;
    part += s;
  }
  const float mean = block_sum(part, red) / (float)C;
  float q = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float d = mean_c[c] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, red) / (float)C + eps);
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(long long)b * C + c] = (mean_c[c] - mean) * rstd * ln_w[c] + ln_b[c];
}

inline unsigned grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 32;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// plain depthwise conv (no LayerNorm) through the register-window kernel; returns 1 when the shape is not specialised
int acb_dwconv7_fast(const void* x, int dtype, const float* w, const float* bias, int flip, void* y, int B, int H, int W, int C, cudaStream_t st) {
  const bool aligned = (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
  if (!((W == 15 || W == 7 || W == 3 || W == 1) && H == W && C % 8 == 0 && C <= 768 && aligned)) return 1;
  return dtype == ACB_F32 ? dispatch_dwconv_w<float>(W, x, w, bias, nullptr, nullptr, 0.0f, y, B, H, C, st, 0, flip)
                          : dispatch_dwconv_w<bf16>(W, x, w, bias, nullptr, nullptr, 0.0f, y, B, H, C, st, 0, flip);
}

extern "C" {

int acb_patchify_nchw(const float* img, int B, int Cin, int H, int W, int p, void* out, int out_dtype, void* stream) {
  ACB_CHECK(img && out && B > 0 && Cin > 0 && p > 0 && H >= p && W >= p, "acb_patchify_nchw: bad arguments");
  const long long n = (long long)B * (H / p) * (W / p) * Cin * p * p;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == ACB_F32)
    patchify_kernel<float><<<grid_for(n), 256, 0, st>>>(img, B, Cin, H, W, p, (float*)out);
  else if (p == 4 && ((uintptr_t)out % 8 == 0))
    patchify4_bf16_kernel<<<grid_for(n / 4), 256, 0, st>>>(img, B, Cin, H, W, (bf16*)out);
  else
    patchify_kernel<bf16><<<grid_for(n), 256, 0, st>>>(img, B, Cin, H, W, p, (bf16*)out);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_dwconv7_ln(const void* x, int dtype, const float* w, const float* b, const float* ln_w, const float* ln_b,
                   float eps, void* y, int B, int H, int W, int C, void* stream) {
  ACB_CHECK(x && y && w && b && ln_w && ln_b && B > 0 && H > 0 && W > 0 && C > 0, "acb_dwconv7_ln: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)ln_w | (uintptr_t)ln_b) % 16 == 0);
  if ((W == 15 || W == 7 || W == 3 || W == 1) && C % 8 == 0 && C <= 768 && aligned) {
    const int rc = dtype == ACB_F32 ? dispatch_dwconv_w<float>(W, x, w, b, ln_w, ln_b, eps, y, B, H, C, st)
                                    : dispatch_dwconv_w<bf16>(W, x, w, b, ln_w, ln_b, eps, y, B, H, C, st);
    if (rc <= 0) return rc;
  }
  const size_t smem = (size_t)8 * W * C * sizeof(float);
  ACB_CHECK(smem <= 200 * 1024, "acb_dwconv7_ln: row tile W*C=%d too large", W * C);
  const unsigned grid = (unsigned)((long long)B * H);
  if (dtype == ACB_F32) {
    auto k = dwconv7_ln_kernel<float>;
    if (smem > 48 * 1024) ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, st>>>((const float*)x, w, b, ln_w, ln_b, eps, (float*)y, H, W, C);
  } else {
    auto k = dwconv7_ln_kernel<bf16>;
    if (smem > 48 * 1024) ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, st>>>((const bf16*)x, w, b, ln_w, ln_b, eps, (bf16*)y, H, W, C);
  }
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_ln_patch2(const void* x, int dtype, const float* ln_w, const float* ln_b, float eps, void* out, int B, int H,
                  int W, int C, void* stream) {
  ACB_CHECK(x && out && ln_w && ln_b && B > 0 && H >= 2 && W >= 2 && C > 0, "acb_ln_patch2: bad arguments");
  const long long pix = (long long)B * (H / 2) * 2 * (W / 2) * 2;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((pix + 7) / 8);
  if (dtype == ACB_BF16 && (C == 96 || C == 192 || C == 384)) {
    constexpr int PPW = 4;
    const unsigned g4 = (unsigned)((pix + 8 * PPW - 1) / (8 * PPW));
    if (C == 96) ln_patch2_stream_kernel<3, PPW><<<g4, 256, 0, st>>>((const bf16*)x, ln_w, ln_b, eps, (bf16*)out, B, H, W);
    else if (C == 192) ln_patch2_stream_kernel<6, PPW><<<g4, 256, 0, st>>>((const bf16*)x, ln_w, ln_b, eps, (bf16*)out, B, H, W);
    else ln_patch2_stream_kernel<12, PPW><<<g4, 256, 0, st>>>((const bf16*)x, ln_w, ln_b, eps, (bf16*)out, B, H, W);
    ACB_LAUNCH_CHECK();
    acb_count_launch();
    return ACB_OK;
  }
  if (dtype == ACB_F32)
    ln_patch2_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ln_w, ln_b, eps, (float*)out, B, H, W, C);
  else
    ln_patch2_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, ln_w, ln_b, eps, (bf16*)out, B, H, W, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_gap_ln(const void* x, int dtype, const float* ln_w, const float* ln_b, float eps, float* out, int B, int HW,
               int C, void* stream) {
  ACB_CHECK(x && out && ln_w && ln_b && B > 0 && HW > 0 && C > 0, "acb_gap_ln: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(C + 40) * sizeof(float);
  if (dtype == ACB_F32)
    gap_ln_kernel<float><<<B, 256, smem, st>>>((const float*)x, ln_w, ln_b, eps, out, HW, C);
  else
    gap_ln_kernel<bf16><<<B, 256, smem, st>>>((const bf16*)x, ln_w, ln_b, eps, out, HW, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
