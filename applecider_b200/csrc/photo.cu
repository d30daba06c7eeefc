// Photometry-transformer kernels: key-padding compaction (varlen packing), fused input embedding
// (Linear(7->D) + Time2Vec + CLS), fused masked varlen attention, CLS read-out.
#include "common.cuh"

namespace {

// ---- compaction ----------------------------------------------------------------------------------
__global__ void count_valid_kernel(const uint8_t* __restrict__ pad, int B, int L, int* __restrict__ cu) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int cnt = 0;
  for (int l = lane; l < L; l += 32) cnt += (pad[(long long)b * L + l] == 0);
  cnt = (int)warp_sum((float)cnt);  // L <= 2^23 so the float sum is exact
  if (lane == 0) cu[b + 1] = cnt + 1;  // + CLS token
  if (b == 0 && lane == 0) cu[0] = 0;
}

// in-place inclusive scan of cu[1..B] by a single block
__global__ void scan_kernel(int* cu, int B, int capacity) {
  __shared__ int sh[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    int v = (i < B) ? cu[i + 1] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < B) cu[i + 1] = min(sh[threadIdx.x] + carry, capacity);  // a too-small caller bound truncates, never overruns
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
}

__global__ void fill_src_kernel(const uint8_t* __restrict__ pad, int B, int L, const int* __restrict__ cu,
                                int* __restrict__ src) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int pos = cu[b];
  const int end = cu[b + 1];  // == pos + 1 + valid count unless the caller's capacity truncated the sequence
  if (lane == 0 && pos < end) src[pos] = -1 - b;
  pos += 1;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const bool valid = (l < L) && (pad[(long long)b * L + l] == 0);
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int at = pos + __popc(m & ((1u << lane) - 1u));
    if (valid && at < end) src[at] = b * L + l;
    pos += __popc(m);
  }
}

// ---- embedding -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) photo_embed_kernel(const float* __restrict__ x, const int* __restrict__ src,
                                                          const int* __restrict__ total_dev, int max_tokens, int D,
                                                          const float* __restrict__ w_in, const float* __restrict__ b_in,
                                                          const float* __restrict__ w0, const float* __restrict__ b0,
                                                          const float* __restrict__ w, const float* __restrict__ bb,
                                                          const float* __restrict__ cls, float te_drop_p,
                                                          AcbSeed te_seed_s, T* __restrict__ out) {
  const unsigned long long te_seed = te_seed_s.get();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float drop_inv = 1.0f / (1.0f - te_drop_p);
  const unsigned drop_thr = (unsigned)(te_drop_p * 4294967296.0);
  const int total = total_dev ? min(*total_dev, max_tokens) : max_tokens;
  if (t >= total) return;
  const int s = src[t];
  T* o = out + (long long)t * D;
  if (s <= ACB_SRC_DEAD) {  // capacity row beyond the packed tokens: zeros (row-local ops keep it inert)
    for (int c = lane; c < D; c += 32) o[c] = from_f<T>(0.0f);
    return;
  }
  if (s < 0) {
    for (int c = lane; c < D; c += 32) o[c] = from_f<T>(cls[c]);
    return;
  }
  float xv = (lane < 7) ? x[(long long)s * 7 + lane] : 0.0f;
  float xr[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) xr[j] = __shfl_sync(0xffffffffu, xv, j);
  const float tt = xr[0];
  for (int c = lane; c < D; c += 32) {
    float h = b_in[c];
#pragma unroll
    for (int j = 0; j < 7; ++j) h = fmaf(w_in[c * 7 + j], xr[j], h);
    float te = (c == 0) ? (w0[0] * tt + b0[0]) : sinf(tt * w[c - 1] + bb[c - 1]);
    if (te_drop_p > 0.0f) te = (te_hash(te_seed, t, c) < drop_thr) ? 0.0f : te * drop_inv;  // MPTModel: F.dropout(te) (:248)
    o[c] = from_f<T>(h + te);
  }
}

// ---- attention -----------------------------------------------------------------------------------
// One CTA per (sequence, head). K/V of the head live in shared memory (fp32); every thread owns
// query rows and runs an exact two-pass softmax (max, then exp/sum/PV) against broadcast K/V reads.
__device__ __forceinline__ unsigned attn_hash(unsigned long long seed, int bh, int i, int j) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)bh << 26) | ((unsigned long long)i << 13) | (unsigned long long)j);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}

template <typename T, int DH>
__global__ void __launch_bounds__(128) attention_varlen_kernel(const T* __restrict__ qkv, const int* __restrict__ cu,
                                                               int n_heads, float drop_p, AcbSeed seed_s,
                                                               T* __restrict__ out) {
  const unsigned long long seed = seed_s.get();
  extern __shared__ float smem[];
  const int b = blockIdx.x, h = blockIdx.y;
  const int t0 = cu[b], n = cu[b + 1] - t0;
  const int D = n_heads * DH;
  float* Ks = smem;
  float* Vs = smem + (size_t)n * DH;
  for (int i = threadIdx.x; i < n * DH; i += blockDim.x) {
    const int r = i / DH, c = i - r * DH;
    const T* row = qkv + (long long)(t0 + r) * 3 * D;
    Ks[i] = to_f<T>(row[D + h * DH + c]);
    Vs[i] = to_f<T>(row[2 * D + h * DH + c]);
  }
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  const float drop_inv = 1.0f / (1.0f - drop_p);
  const unsigned drop_thr = (unsigned)(drop_p * 4294967296.0);
  const int bh = b * n_heads + h;
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const T* qrow = qkv + (long long)(t0 + r) * 3 * D + h * DH;
    float q[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) q[c] = to_f<T>(qrow[c]) * scale;
    float m = -INFINITY;
    for (int j = 0; j < n; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(Ks + j * DH);
      float s = 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 k4 = kp[c4];
        s = fmaf(q[c4 * 4 + 0], k4.x, s);
        s = fmaf(q[c4 * 4 + 1], k4.y, s);
        s = fmaf(q[c4 * 4 + 2], k4.z, s);
        s = fmaf(q[c4 * 4 + 3], k4.w, s);
      }
      m = fmaxf(m, s);
    }
    float l = 0.0f, acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.0f;
    for (int j = 0; j < n; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(Ks + j * DH);
      const float4* vp = reinterpret_cast<const float4*>(Vs + j * DH);
      float s = 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 k4 = kp[c4];
        s = fmaf(q[c4 * 4 + 0], k4.x, s);
        s = fmaf(q[c4 * 4 + 1], k4.y, s);
        s = fmaf(q[c4 * 4 + 2], k4.z, s);
        s = fmaf(q[c4 * 4 + 3], k4.w, s);
      }
      float p = expf(s - m);
      l += p;
      if (drop_p > 0.0f) p = attn_hash(seed, bh, r, j) >= drop_thr ? p * drop_inv : 0.0f;
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 v4 = vp[c4];
        acc[c4 * 4 + 0] = fmaf(p, v4.x, acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(p, v4.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(p, v4.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(p, v4.w, acc[c4 * 4 + 3]);
      }
    }
    const float inv = 1.0f / l;
    T* orow = out + (long long)(t0 + r) * D + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) orow[c] = from_f<T>(acc[c] * inv);
  }
}

template <typename T>
__global__ void gather_cls_kernel(const T* __restrict__ x, const int* __restrict__ cu, int B, int D, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D) return;
  const int c = (int)(i % D);
  const int b = (int)(i / D);
  out[i] = to_f<T>(x[(long long)cu[b] * D + c]);
}

}  // namespace

extern "C" {

int acb_photo_compact(const uint8_t* pad, int B, int L, int capacity, int* cu_seqlens, int* src_idx, void* stream) {
  ACB_CHECK(pad && cu_seqlens && src_idx && B > 0 && L > 0, "acb_photo_compact: bad arguments");
  ACB_CHECK((long long)B * (L + 1) < (1LL << 30), "acb_photo_compact: B*(L+1) = %lld tokens exceed the 2^30 index range", (long long)B * (L + 1));
  if (capacity <= 0 || capacity > B * (L + 1)) capacity = B * (L + 1);
  cudaStream_t st = (cudaStream_t)stream;
  // every entry the fill below does not reach (rows past cu[B] of the B*(L+1) capacity) reads as "dead": 0x80808080 < ACB_SRC_DEAD
  ACB_CUDA(cudaMemsetAsync(src_idx, 0x80, (size_t)B * (L + 1) * sizeof(int), st));
  count_valid_kernel<<<cdiv(B, 8), 256, 0, st>>>(pad, B, L, cu_seqlens);
  ACB_LAUNCH_CHECK();
  scan_kernel<<<1, 1024, 0, st>>>(cu_seqlens, B, capacity);
  ACB_LAUNCH_CHECK();
  fill_src_kernel<<<cdiv(B, 8), 256, 0, st>>>(pad, B, L, cu_seqlens, src_idx);
  ACB_LAUNCH_CHECK();
  acb_count_launch(3);
  return ACB_OK;
}

int acb_photo_embed(const float* x, const int* src_idx, const int* total_dev, int max_tokens, int D, const float* w_in,
                    const float* b_in, const float* w0, const float* b0, const float* w, const float* b,
                    const float* cls_tok, float te_drop_p, long long te_seed, void* out, int out_dtype, void* stream) {
  ACB_CHECK(x && src_idx && out && max_tokens >= 0 && D > 1, "acb_photo_embed: bad arguments");
  ACB_CHECK(te_drop_p >= 0.0f && te_drop_p < 1.0f, "acb_photo_embed: dropout %g out of range", (double)te_drop_p);
  if (max_tokens == 0) return ACB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv(max_tokens, 8);
  if (out_dtype == ACB_F32)
    photo_embed_kernel<float><<<grid, 256, 0, st>>>(x, src_idx, total_dev, max_tokens, D, w_in, b_in, w0, b0, w, b, cls_tok, te_drop_p, acb_seed(te_seed), (float*)out);
  else
    photo_embed_kernel<bf16><<<grid, 256, 0, st>>>(x, src_idx, total_dev, max_tokens, D, w_in, b_in, w0, b0, w, b, cls_tok, te_drop_p, acb_seed(te_seed), (bf16*)out);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_attention_varlen(const void* qkv, int dtype, const int* cu_seqlens, int B, int n_heads, int dh, int max_seqlen,
                         float drop_p, long long seed, void* out, void* stream) {
  ACB_CHECK(qkv && cu_seqlens && out && B > 0 && n_heads > 0, "acb_attention_varlen: bad arguments");
  ACB_CHECK(dh == 16, "acb_attention_varlen: head dim %d unsupported (16 only)", dh);
  ACB_CHECK(drop_p >= 0.0f && drop_p < 1.0f && max_seqlen < 8192, "acb_attention_varlen: bad dropout / sequence length");
  const size_t smem = (size_t)max_seqlen * dh * 2 * sizeof(float);
  ACB_CHECK(smem <= 200 * 1024, "acb_attention_varlen: max_seqlen %d too long for the shared-memory K/V tile", max_seqlen);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(B, n_heads);
  if (dtype == ACB_F32) {
    auto k = attention_varlen_kernel<float, 16>;
    if (smem > 48 * 1024) ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 128, smem, st>>>((const float*)qkv, cu_seqlens, n_heads, drop_p, acb_seed(seed), (float*)out);
  } else {
    auto k = attention_varlen_kernel<bf16, 16>;
    if (smem > 48 * 1024) ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 128, smem, st>>>((const bf16*)qkv, cu_seqlens, n_heads, drop_p, acb_seed(seed), (bf16*)out);
  }
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_gather_cls(const void* x, int dtype, const int* cu_seqlens, int B, int D, float* out, void* stream) {
  ACB_CHECK(x && cu_seqlens && out && B > 0 && D > 0, "acb_gather_cls: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv((long long)B * D, 256);
  if (dtype == ACB_F32)
    gather_cls_kernel<float><<<grid, 256, 0, st>>>((const float*)x, cu_seqlens, B, D, out);
  else
    gather_cls_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, cu_seqlens, B, D, out);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
