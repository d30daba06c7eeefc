// Masked-event pre-training (reference: MPTModel, models/HyraxBaselineCLS.py:194-319).
//   acb_mpt_mask          band-balanced token sampling + in-place zeroing of channels 2:7 (_mask_batch :283-319)
//   acb_mpt_loss_fwd_bwd  the three masked-token losses, their PRODUCT (:278, sic) and d loss / d head outputs
// No host synchronisation anywhere: counts, sums and the product rule stay on the device.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned mpt_hash(unsigned long long seed, int b, int i, int stream_id) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)b << 24) | ((unsigned long long)i << 2) | (unsigned long long)stream_id);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return (unsigned)v;
}

// One CTA per light curve.  Every valid token gets two random keys; "take the first `take` of a random
// permutation" == "take the `take` smallest keys", evaluated by rank counting (L <= a few hundred).
__global__ void __launch_bounds__(128) mpt_mask_kernel(float* __restrict__ x, const uint8_t* __restrict__ pad, int L, double mask_p,
                                                       AcbSeed seed_s, uint8_t* __restrict__ masked) {
  const unsigned long long seed = seed_s.get();
  extern __shared__ unsigned sm_u[];
  unsigned* key = sm_u;                                  // [L]
  signed char* band = reinterpret_cast<signed char*>(key + L);  // [L]  -1 = padded
  signed char* sel = band + L;                           // [L]
  __shared__ int cnt[4];                                 // per band, [3] = valid
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 4) cnt[tid] = 0;
  __syncthreads();
  for (int i = tid; i < L; i += blockDim.x) {
    int bd = -1;
    if (!pad[(long long)b * L + i]) {
      const float* r = x + ((long long)b * L + i) * 7 + 4;
      bd = 0;  // argmax with first-index tie break
      float best = r[0];
      if (r[1] > best) { best = r[1]; bd = 1; }
      if (r[2] > best) { bd = 2; }
      atomicAdd(&cnt[bd], 1);
      atomicAdd(&cnt[3], 1);
    }
    band[i] = (signed char)bd;
    sel[i] = 0;
    key[i] = mpt_hash(seed, b, i, 0);
  }
  __syncthreads();
  const int n_valid = cnt[3];
  int k = (int)((double)n_valid * mask_p);  // int(len(valid) * MASK_P), python float arithmetic
  if (k < 3) k = 3;
  const int num_each = k / 3, extras = k - 3 * num_each;
  for (int i = tid; i < L; i += blockDim.x) {
    const int bd = band[i];
    if (bd < 0) continue;
    const int take = min(cnt[bd], num_each);
    const unsigned ki = key[i];
    int rank = 0;
    for (int j = 0; j < L; ++j)
      if (band[j] == bd && (key[j] < ki || (key[j] == ki && j < i))) ++rank;
    if (rank < take) sel[i] = 1;
  }
  __syncthreads();
  if (extras > 0) {  // uniformly from the valid tokens not taken yet
    for (int i = tid; i < L; i += blockDim.x) {
      if (band[i] < 0 || sel[i]) continue;
      const unsigned ki = mpt_hash(seed, b, i, 1);
      int rank = 0;
      for (int j = 0; j < L; ++j) {
        if (band[j] < 0 || sel[j] == 1) continue;
        const unsigned kj = mpt_hash(seed, b, j, 1);
        if (kj < ki || (kj == ki && j < i)) ++rank;
      }
      if (rank < extras) sel[i] = 2;
    }
    __syncthreads();
  }
  for (int i = tid; i < L; i += blockDim.x) {
    const bool m = sel[i] != 0;
    masked[(long long)b * L + i] = m ? 1 : 0;
    if (m) {
      float* r = x + ((long long)b * L + i) * 7;
#pragma unroll
      for (int c = 2; c < 7; ++c) r[c] = 0.0f;
    }
  }
}

// targets exactly as the reference reads them AFTER the in-place masking (true_f and the band one-hot of masked
// tokens are therefore 0 / class 0 -- HyraxBaselineCLS.py:264-272 read `data`, which _mask_batch has modified)
struct MptTok {
  float f_hat, b_hat[3], dt_hat, tf, dt_gt;
  int tb;
};
__device__ __forceinline__ MptTok mpt_load(const void* pred, int pdt, long long t, const float* x, int s, int L) {
  MptTok k;
  k.f_hat = ld_any(pred, t * 5 + 0, pdt);
  k.b_hat[0] = ld_any(pred, t * 5 + 1, pdt);
  k.b_hat[1] = ld_any(pred, t * 5 + 2, pdt);
  k.b_hat[2] = ld_any(pred, t * 5 + 3, pdt);
  k.dt_hat = ld_any(pred, t * 5 + 4, pdt);
  const float* r = x + (long long)s * 7;
  k.tf = r[2];
  k.tb = 0;
  float best = r[4];
  if (r[5] > best) { best = r[5]; k.tb = 1; }
  if (r[6] > best) { k.tb = 2; }
  const int i = s % L;
  k.dt_gt = (i + 1 < L) ? r[7 + 1] : 0.0f;  // roll(data[...,1], -1) with the last column zeroed
  return k;
}

__global__ void __launch_bounds__(256) mpt_loss_sum_kernel(const void* pred, int pdt, const int* __restrict__ src, int T,
                                                           const float* __restrict__ x, const uint8_t* __restrict__ masked, int L,
                                                           float* __restrict__ ws) {
  float sf = 0.f, sb = 0.f, sd = 0.f, sc = 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
    const int s = src[t];
    if (s < 0 || !masked[s]) continue;
    const MptTok k = mpt_load(pred, pdt, t, x, s, L);
    const float df = k.f_hat - k.tf, dd = k.dt_hat - k.dt_gt;
    const float mx = fmaxf(k.b_hat[0], fmaxf(k.b_hat[1], k.b_hat[2]));
    const float lse = mx + logf(expf(k.b_hat[0] - mx) + expf(k.b_hat[1] - mx) + expf(k.b_hat[2] - mx));
    sf += df * df;
    sb += lse - k.b_hat[k.tb];
    sd += dd * dd;
    sc += 1.0f;
  }
  __shared__ float sh[33];
  sf = block_sum(sf, sh); sb = block_sum(sb, sh); sd = block_sum(sd, sh); sc = block_sum(sc, sh);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 0, sf); atomicAdd(ws + 1, sb); atomicAdd(ws + 2, sd); atomicAdd(ws + 3, sc);
  }
}

__global__ void __launch_bounds__(256) mpt_loss_grad_kernel(const void* pred, int pdt, const int* __restrict__ src, int T,
                                                            const float* __restrict__ x, const uint8_t* __restrict__ masked, int L,
                                                            const float* __restrict__ ws, float lam, float* __restrict__ losses,
                                                            float* __restrict__ dpred) {
  const float cnt = ws[3];
  const float Lf = ws[0] / cnt, Lb = ws[1] / cnt, Ld = ws[2] / cnt;  // mean over the masked tokens (0/0 = nan like torch)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    losses[0] = lam * Lf * Lb * Ld;
    losses[1] = Lf; losses[2] = Lb; losses[3] = Ld;
  }
  if (!dpred) return;
  const float gf = lam * Lb * Ld / cnt, gb = lam * Lf * Ld / cnt, gd = lam * Lf * Lb / cnt;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
    const int s = src[t];
    float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (s >= 0 && masked[s]) {
      const MptTok k = mpt_load(pred, pdt, t, x, s, L);
      g[0] = 2.0f * (k.f_hat - k.tf) * gf;
      const float mx = fmaxf(k.b_hat[0], fmaxf(k.b_hat[1], k.b_hat[2]));
      const float e0 = expf(k.b_hat[0] - mx), e1 = expf(k.b_hat[1] - mx), e2 = expf(k.b_hat[2] - mx);
      const float inv = 1.0f / (e0 + e1 + e2);
      g[1] = (e0 * inv - (k.tb == 0 ? 1.f : 0.f)) * gb;
      g[2] = (e1 * inv - (k.tb == 1 ? 1.f : 0.f)) * gb;
      g[3] = (e2 * inv - (k.tb == 2 ? 1.f : 0.f)) * gb;
      g[4] = 2.0f * (k.dt_hat - k.dt_gt) * gd;
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) dpred[t * 5 + c] = g[c];
  }
}

}  // namespace

extern "C" {

int acb_mpt_mask(float* x, const uint8_t* pad, int B, int L, double mask_p, long long seed, uint8_t* masked, void* stream) {
  ACB_CHECK(x && pad && masked && B > 0 && L > 0, "acb_mpt_mask: bad arguments");
  ACB_CHECK(mask_p >= 0.0 && mask_p <= 1.0 && L <= 8192, "acb_mpt_mask: mask_p %g / L %d out of range", mask_p, L);
  const size_t smem = (size_t)L * 4 + 2 * (size_t)L + 16;
  mpt_mask_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(x, pad, L, mask_p, acb_seed(seed), masked);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_mpt_loss_fwd_bwd(const void* pred, int pred_dtype, const int* src_idx, int T, const float* x, const uint8_t* masked, int L,
                         float lambda_f, float lambda_b, float lambda_dt, float* losses, float* dpred, float* workspace, void* stream) {
  ACB_CHECK(pred && src_idx && x && masked && losses && workspace && T > 0 && L > 0, "acb_mpt_loss_fwd_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(workspace, 0, 4 * sizeof(float), st));
  const int grid = min(cdiv(T, 256), 148 * 8);
  mpt_loss_sum_kernel<<<grid, 256, 0, st>>>(pred, pred_dtype, src_idx, T, x, masked, L, workspace);
  ACB_LAUNCH_CHECK();
  mpt_loss_grad_kernel<<<grid, 256, 0, st>>>(pred, pred_dtype, src_idx, T, x, masked, L, workspace, lambda_f * lambda_b * lambda_dt, losses,
                                             dpred);
  ACB_LAUNCH_CHECK();
  acb_count_launch(2);
  return ACB_OK;
}

}  // extern "C"
