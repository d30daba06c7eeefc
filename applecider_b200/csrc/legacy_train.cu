// Training kernels of the legacy "variant B" spectra encoder (_archive/notebooks/brew_cider.py:585-636): BatchNorm1d over
// channels-last [rows, C] activations (batch statistics + running-statistics update in training, running statistics in
// eval), fused  out = act(res + BN(y))  with its backward, and the max / avg / min tri-pool backward.
// All HBM-bound elementwise / column-reduction work: one pass per tensor, fp32 statistics.
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_grad_any(float pre, int act) {
  switch (act) {
    case ACB_ACT_RELU: return pre > 0.0f ? 1.0f : 0.0f;
    case ACB_ACT_GELU: return gelu_erf_grad(pre);
    case ACB_ACT_TANH: { const float t = tanhf(pre); return 1.0f - t * t; }
    case ACB_ACT_SIGMOID: { const float s = sigmoidf_(pre); return s * (1.0f - s); }
    default: return 1.0f;
  }
}

// per-channel: statistics -> (scale, shift, mean, rstd); training also blends the running statistics (momentum, unbiased var)
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, long long M, int C,
                                   const float* __restrict__ w, const float* __restrict__ b, float eps, float momentum, int training,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const double m = (double)sum[c] / (double)M;
    double v = (double)sumsq[c] / (double)M - m * m;  // biased (what normalises the batch)
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      const double unb = M > 1 ? v * (double)M / (double)(M - 1) : v;
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float s = w[c] * rstd;
  scale[c] = s;
  shift[c] = b[c] - mean * s;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
}

// column sums of y and y^2 (fp32 accumulation, one atomic pair per thread)
__global__ void __launch_bounds__(256) bn_stats_kernel(const void* __restrict__ y, int y_dt, long long rows, int C, long long stride_elems,
                                                       float* __restrict__ sums) {
  const long long total = rows * C;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= stride_elems) return;
  const int c = (int)(i0 % C);  // stride_elems % C == 0: this thread always sees channel c
  float s = 0.0f, q = 0.0f;
  for (long long i = i0; i < total; i += stride_elems) {
    const float v = ld_any(y, i, y_dt);
    s += v;
    q = fmaf(v, v, q);
  }
  atomicAdd(sums + c, s);
  atomicAdd(sums + C + c, q);
}

__global__ void __launch_bounds__(256) affine_res_act_kernel(const void* __restrict__ y, int y_dt, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const void* __restrict__ res, int res_dt,
                                                             int act, void* __restrict__ out, int out_dt, long long total, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float v = fmaf(scale[c], ld_any(y, i, y_dt), shift[c]);
    if (res) v += ld_any(res, i, res_dt);
    st_any(out, i, out_dt, apply_act(v, act));
  }
}

// dpre = dout * act'(pre) (pre recomputed), column sums of dpre and dpre * xhat
__global__ void __launch_bounds__(256) affine_res_act_bwd_kernel(const void* __restrict__ y, int y_dt, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, const void* __restrict__ res, int res_dt,
                                                                 int act, const void* __restrict__ dout, int dout_dt,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 void* __restrict__ dpre, int dpre_dt, float* __restrict__ sums,
                                                                 long long rows, int C, long long stride_elems) {
  const long long total = rows * C;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= stride_elems) return;
  const int c = (int)(i0 % C);
  const float sc = scale[c], sh = shift[c], mu = mean[c], rs = rstd[c];
  float s1 = 0.0f, s2 = 0.0f;
  for (long long i = i0; i < total; i += stride_elems) {
    const float yv = ld_any(y, i, y_dt);
    float pre = fmaf(sc, yv, sh);
    if (res) pre += ld_any(res, i, res_dt);
    const float g = ld_any(dout, i, dout_dt) * act_grad_any(pre, act);
    st_any(dpre, i, dpre_dt, g);
    s1 += g;
    s2 = fmaf(g, (yv - mu) * rs, s2);
  }
  atomicAdd(sums + c, s1);
  atomicAdd(sums + C + c, s2);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const void* __restrict__ y, int y_dt, const void* __restrict__ dpre, int dpre_dt,
                                                           const float* __restrict__ scale, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ sums, int training,
                                                           void* __restrict__ dy, int dy_dt, long long rows, int C) {
  const long long total = rows * C;
  const float invM = 1.0f / (float)rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float g = ld_any(dpre, i, dpre_dt);
    if (training) {
      const float xhat = (ld_any(y, i, y_dt) - mean[c]) * rstd[c];
      g = g - sums[c] * invM - xhat * sums[C + c] * invM;
    }
    st_any(dy, i, dy_dt, scale[c] * g);
  }
}

__global__ void tripool4_any_kernel(const void* __restrict__ x, int x_dt, void* __restrict__ y, int y_dt, int B, int L, int C) {
  const int Lo = L / 4;
  const long long total = (long long)B * Lo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int lo = (int)(t % Lo);
    const long long b = t / Lo;
    const long long p = (b * L + 4LL * lo) * C + c;
    const float v0 = ld_any(x, p, x_dt), v1 = ld_any(x, p + C, x_dt), v2 = ld_any(x, p + 2LL * C, x_dt), v3 = ld_any(x, p + 3LL * C, x_dt);
    const long long o = (b * Lo + lo) * (3LL * C) + c;
    st_any(y, o, y_dt, fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));
    st_any(y, o + C, y_dt, (((v0 + v1) + v2) + v3) * 0.25f);
    st_any(y, o + 2LL * C, y_dt, fminf(fminf(v0, v1), fminf(v2, v3)));
  }
}

// dx[l] = dmax * [l == first argmax] + davg / 4 + dmin * [l == first argmin]   (torch routes ties to the first index)
__global__ void tripool4_bwd_kernel(const void* __restrict__ x, int x_dt, const void* __restrict__ dy, int dy_dt, void* __restrict__ dx,
                                    int dx_dt, int B, int L, int C) {
  const int Lo = L / 4;
  const long long total = (long long)B * Lo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long t = i / C;
    const int lo = (int)(t % Lo);
    const long long b = t / Lo;
    const long long p = (b * L + 4LL * lo) * C + c;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_any(x, p + (long long)k * C, x_dt);
    int imax = 0, imin = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      if (v[k] > v[imax]) imax = k;
      if (v[k] < v[imin]) imin = k;
    }
    const long long o = (b * Lo + lo) * (3LL * C) + c;
    const float gmax = ld_any(dy, o, dy_dt), gavg = ld_any(dy, o + C, dy_dt) * 0.25f, gmin = ld_any(dy, o + 2LL * C, dy_dt);
#pragma unroll
    for (int k = 0; k < 4; ++k) st_any(dx, p + (long long)k * C, dx_dt, gavg + (k == imax ? gmax : 0.0f) + (k == imin ? gmin : 0.0f));
  }
  // positions past 4 * (L / 4) receive no gradient
  const int tail = L - 4 * Lo;
  if (tail > 0) {
    const long long tt = (long long)B * tail * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tt; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(i % C);
      const long long t = i / C;
      const int l = 4 * Lo + (int)(t % tail);
      const long long b = t / tail;
      st_any(dx, (b * L + l) * C + c, dx_dt, 0.0f);
    }
  }
}

inline unsigned ew_grid(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148LL * 16); }

// grid whose total thread count is a multiple of C (every thread then owns one channel in a grid-stride loop)
inline void col_grid(long long rows, int C, unsigned* grid, long long* stride) {
  long long want = std::min<long long>(rows * C, 148LL * 8 * 256);
  long long s = (want + C - 1) / C * C;
  *stride = s;
  *grid = (unsigned)((s + 255) / 256);
}

}  // namespace

extern "C" {

int acb_bn_stats(const void* y, int y_dtype, long long rows, int C, float* sums, void* stream) {
  ACB_CHECK(y && sums && rows > 0 && C > 0, "acb_bn_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  unsigned grid; long long stride;
  col_grid(rows, C, &grid, &stride);
  bn_stats_kernel<<<grid, 256, 0, st>>>(y, y_dtype, rows, C, stride, sums);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_bn_finalize(const float* sums, long long rows, int C, const float* w, const float* b, float eps, float momentum, int training,
                    float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* rstd, void* stream) {
  ACB_CHECK(w && b && scale && shift && mean && rstd && C > 0 && rows > 0, "acb_bn_finalize: bad arguments");
  ACB_CHECK(training ? sums != nullptr : (running_mean && running_var), "acb_bn_finalize: missing statistics");
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, sums ? sums + C : nullptr, rows, C, w, b, eps, momentum, training,
                                                                    running_mean, running_var, scale, shift, mean, rstd);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_affine_res_act(const void* y, int y_dtype, const float* scale, const float* shift, const void* res, int res_dtype, int act,
                       void* out, int out_dtype, long long rows, int C, void* stream) {
  ACB_CHECK(y && scale && shift && out && rows > 0 && C > 0, "acb_affine_res_act: bad arguments");
  affine_res_act_kernel<<<ew_grid(rows * C), 256, 0, (cudaStream_t)stream>>>(y, y_dtype, scale, shift, res, res_dtype, act, out, out_dtype,
                                                                            rows * C, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_affine_res_act_bwd(const void* y, int y_dtype, const float* scale, const float* shift, const void* res, int res_dtype, int act,
                           const void* dout, int dout_dtype, const float* mean, const float* rstd, void* dpre, int dpre_dtype,
                           float* sums, long long rows, int C, void* stream) {
  ACB_CHECK(y && scale && shift && dout && mean && rstd && dpre && sums && rows > 0 && C > 0, "acb_affine_res_act_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(float), st));
  unsigned grid; long long stride;
  col_grid(rows, C, &grid, &stride);
  affine_res_act_bwd_kernel<<<grid, 256, 0, st>>>(y, y_dtype, scale, shift, res, res_dtype, act, dout, dout_dtype, mean, rstd, dpre, dpre_dtype,
                                                  sums, rows, C, stride);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_bn_bwd_apply(const void* y, int y_dtype, const void* dpre, int dpre_dtype, const float* scale, const float* mean, const float* rstd,
                     const float* sums, int training, void* dy, int dy_dtype, long long rows, int C, void* stream) {
  ACB_CHECK(y && dpre && scale && mean && rstd && sums && dy && rows > 0 && C > 0, "acb_bn_bwd_apply: bad arguments");
  bn_bwd_apply_kernel<<<ew_grid(rows * C), 256, 0, (cudaStream_t)stream>>>(y, y_dtype, dpre, dpre_dtype, scale, mean, rstd, sums, training, dy,
                                                                          dy_dtype, rows, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_tripool4(const void* x, int x_dtype, void* y, int y_dtype, int B, int L, int C, void* stream) {
  ACB_CHECK(x && y && B > 0 && L >= 4 && C > 0, "acb_tripool4: bad arguments");
  tripool4_any_kernel<<<ew_grid((long long)B * (L / 4) * C), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, y_dtype, B, L, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_tripool4_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype, int B, int L, int C, void* stream) {
  ACB_CHECK(x && dy && dx && B > 0 && L >= 4 && C > 0, "acb_tripool4_bwd: bad arguments");
  tripool4_bwd_kernel<<<ew_grid((long long)B * (L / 4) * C), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, dy, dy_dtype, dx, dx_dtype, B, L, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
