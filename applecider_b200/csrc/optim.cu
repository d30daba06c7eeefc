// Fused optimiser step over flat fp32 buffers: gradient-norm clip + Adam / AdamW with per-group
// hyper-parameters in ONE pass, optionally refreshing the bf16 shadow copy of the weights that the tcgen05
// GEMMs read.  Reference call sites: HyraxBaselineCLS.py:108-120 (clip 1.0 + Adam), :228 (AdamW),
// astrominn.py:151-218,311-326 (11-group AdamW, eps 5e-10), brew_cider.py:1211 (Adam, L2 weight decay).
// HBM-bound: 16 B read + 12 B (+2 B) written per parameter; float4 accesses, grid = a multiple of the SM count.
#include "common.cuh"

namespace {

constexpr int ADAM_MAX_GROUPS = 16;

struct AdamGroups {
  int n;
  long long end[ADAM_MAX_GROUPS];  // exclusive end offset of every group in the flat buffers
  float lr[ADAM_MAX_GROUPS], b1[ADAM_MAX_GROUPS], b2[ADAM_MAX_GROUPS], eps[ADAM_MAX_GROUPS], wd[ADAM_MAX_GROUPS];
  float step_size[ADAM_MAX_GROUPS];   // lr / (1 - b1^t)
  float inv_bc2_sqrt[ADAM_MAX_GROUPS];  // 1 / sqrt(1 - b2^t)
  int decoupled[ADAM_MAX_GROUPS];
};

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, bf16* __restrict__ p16, long long n4,
                                                        const __grid_constant__ AdamGroups G, const float* __restrict__ gnorm_sq,
                                                        float max_norm, float grad_scale, const int* __restrict__ step_dev) {
  // bias corrections: from the host step count (baked into G) or, when the step lives on the device (CUDA-graph replay: the
  // graph increments *step_dev itself), recomputed here once per block in double precision like the host path
  __shared__ float s_step_size[ADAM_MAX_GROUPS], s_inv_bc2[ADAM_MAX_GROUPS];
  if (threadIdx.x < G.n) {
    const int gi = threadIdx.x;
    if (step_dev) {
      const double t = (double)max(*step_dev, 1);
      s_step_size[gi] = (float)((double)G.lr[gi] / (1.0 - pow((double)G.b1[gi], t)));
      s_inv_bc2[gi] = (float)(1.0 / sqrt(1.0 - pow((double)G.b2[gi], t)));
    } else {
      s_step_size[gi] = G.step_size[gi];
      s_inv_bc2[gi] = G.inv_bc2_sqrt[gi];
    }
  }
  __syncthreads();
  float coef = grad_scale;
  if (gnorm_sq) coef *= fminf(1.0f, max_norm / (sqrtf(*gnorm_sq) * fabsf(grad_scale) + 1e-6f));  // clip_grad_norm_ arithmetic
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
    const long long i0 = q * 4;
    int gi = 0;
    while (gi < G.n - 1 && i0 >= G.end[gi]) ++gi;
    float4 P = reinterpret_cast<float4*>(p)[q];
    const float4 Gr = reinterpret_cast<const float4*>(g)[q];
    float4 M = reinterpret_cast<float4*>(m)[q];
    float4 V = reinterpret_cast<float4*>(v)[q];
    float* pe = reinterpret_cast<float*>(&P);
    const float* ge = reinterpret_cast<const float*>(&Gr);
    float* me = reinterpret_cast<float*>(&M);
    float* ve = reinterpret_cast<float*>(&V);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      while (gi < G.n - 1 && i0 + e >= G.end[gi]) ++gi;
      float grad = ge[e] * coef;
      float w = pe[e];
      if (G.decoupled[gi]) w *= 1.0f - G.lr[gi] * G.wd[gi];  // AdamW
      else grad = fmaf(G.wd[gi], w, grad);                      // Adam: L2 term joins the gradient
      const float mm = fmaf(G.b1[gi], me[e], (1.0f - G.b1[gi]) * grad);
      const float vv = fmaf(G.b2[gi], ve[e], (1.0f - G.b2[gi]) * grad * grad);
      const float denom = sqrtf(vv) * s_inv_bc2[gi] + G.eps[gi];
      pe[e] = w - s_step_size[gi] * (mm / denom);
      me[e] = mm;
      ve[e] = vv;
    }
    reinterpret_cast<float4*>(p)[q] = P;
    reinterpret_cast<float4*>(m)[q] = M;
    reinterpret_cast<float4*>(v)[q] = V;
    if (p16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pe[0], pe[1]), hi = __floats2bfloat162_rn(pe[2], pe[3]);
      uint2 o;
      o.x = *reinterpret_cast<unsigned*>(&lo);
      o.y = *reinterpret_cast<unsigned*>(&hi);
      reinterpret_cast<uint2*>(p16)[q] = o;
    }
  }
}

}  // namespace

extern "C" {

int acb_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, int n_groups, const long long* group_end,
                  const float* hyper, int step, const int* step_dev, const float* gnorm_sq, float max_norm, float grad_scale,
                  void* stream) {
  ACB_CHECK(p && g && m && v && group_end && hyper && n > 0 && step >= 1, "acb_adam_step: bad arguments");
  ACB_CHECK(n_groups >= 1 && n_groups <= ADAM_MAX_GROUPS, "acb_adam_step: %d parameter groups (max %d)", n_groups, ADAM_MAX_GROUPS);
  ACB_CHECK(n % 4 == 0 && (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && (((uintptr_t)p_bf16) & 7) == 0,
            "acb_adam_step: flat buffers must be 16-byte aligned with a length that is a multiple of 4");
  AdamGroups G;
  G.n = n_groups;
  long long prev = 0;
  for (int i = 0; i < n_groups; ++i) {
    const float* h = hyper + 6 * i;  // lr, beta1, beta2, eps, weight_decay, decoupled
    ACB_CHECK(group_end[i] >= prev && group_end[i] <= n, "acb_adam_step: group ends must be non-decreasing and <= n");
    prev = group_end[i];
    G.end[i] = group_end[i];
    G.lr[i] = h[0]; G.b1[i] = h[1]; G.b2[i] = h[2]; G.eps[i] = h[3]; G.wd[i] = h[4];
    G.decoupled[i] = h[5] != 0.0f;
    const double bc1 = 1.0 - pow((double)h[1], (double)step), bc2 = 1.0 - pow((double)h[2], (double)step);
    G.step_size[i] = (float)((double)h[0] / bc1);
    G.inv_bc2_sqrt[i] = (float)(1.0 / sqrt(bc2));
  }
  G.end[n_groups - 1] = n;  // trailing alignment padding belongs to the last group (its gradient is zero)
  const long long n4 = n / 4;
  const int grid = (int)std::min<long long>((n4 + 255) / 256, 148LL * 8);
  adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_bf16, n4, G, gnorm_sq, max_norm, grad_scale, step_dev);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
