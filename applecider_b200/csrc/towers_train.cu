// Training path of the gated residual towers / MoE experts (ResidualTowerBlock, astrominn.py:44-64): a whole GROUP of
// towers (the 8 metadata towers, or the 4 experts) runs as ONE forward launch and ONE backward launch
// (blockIdx.y = tower) instead of ~25 tiny GEMM / LayerNorm / activation / dropout launches per tower.
//   forward : [a = W0 x + b0 | a given] -> s = gelu(a) -> shared LN statistics -> two affine+dropout heads ->
//             y = (W1 n1 + b1) * sigmoid(W2 n2 + b2) + skip(x);  a (pre-GELU) is the only tensor saved.
//   backward: recomputes the row from a, produces d a (expert mode), d x (optional) and ALL parameter gradients;
//             outer-product gradients are formed per 32-row CTA tile from shared memory, vector gradients in registers.
// fp32 throughout (the towers are a few KB of weights and HBM/latency bound).
#include "common.cuh"

namespace {

constexpr int TG_MAX_TOWERS = 8;
constexpr int TG_MAX_IN = 512;
constexpr int TG_MAX_HID = 256;
constexpr int TG_HPL = TG_MAX_HID / 32;  // hidden values per lane
constexpr int TG_ROWS = 32;              // rows per backward CTA
constexpr int TG_BWD_WARPS = 8;

struct TowerT {
  const int* cols;
  int in_dim, hid, out_dim, y_off, a_off, pad_;
  const float *W0, *b0, *ln1w, *ln1b, *W1, *b1, *ln2w, *ln2b, *W2, *b2, *Ws, *bs;
  float *gW0, *gb0, *gln1w, *gln1b, *gW1, *gb1, *gln2w, *gln2b, *gW2, *gb2, *gWs, *gbs;
};
struct TowerGroup {
  int n;
  TowerT t[TG_MAX_TOWERS];
};

__device__ __forceinline__ float tw_keep(unsigned long long seed, int tower, long long row, int h, int which, unsigned thr, float inv) {
  unsigned long long v = seed * 0x9E3779B97F4A7C15ULL + (((unsigned long long)row << 16) | ((unsigned long long)tower << 9) | ((unsigned long long)which << 8) | (unsigned long long)h);
  v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
  return ((unsigned)v < thr) ? 0.0f : inv;
}

// ---- forward (one warp per row) ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tower_group_fwd_kernel(const float* __restrict__ X, int ldx, int rows, const __grid_constant__ TowerGroup G,
                                                              float* __restrict__ Y, int ldy, float* __restrict__ A, int lda, float drop_p,
                                                              AcbSeed seed_s) {
  const unsigned long long seed = seed_s.get();
  __shared__ float xs[4][TG_MAX_IN];
  __shared__ float n1s[4][TG_MAX_HID];
  __shared__ float n2s[4][TG_MAX_HID];
  const TowerT& p = G.t[blockIdx.y];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 4 + wid;
  if (row >= rows) return;
  float* x = xs[wid];
  float* n1 = n1s[wid];
  float* n2 = n2s[wid];
  for (int i = lane; i < p.in_dim; i += 32) x[i] = X[row * ldx + (p.cols ? p.cols[i] : i)];
  __syncwarp();
  float sum = 0.0f;
  for (int h = lane; h < p.hid; h += 32) {
    float a;
    if (p.W0) {
      const float* wr = p.W0 + (long long)h * p.in_dim;
      a = p.b0[h];
      for (int i = 0; i < p.in_dim; ++i) a = fmaf(__ldg(wr + i), x[i], a);
      A[row * lda + p.a_off + h] = a;
    } else {
      a = A[row * lda + p.a_off + h];
    }
    const float s = gelu_erf(a);
    n1[h] = s;
    sum += s;
  }
  const float mean = warp_sum(sum) / (float)p.hid;
  float q = 0.0f;
  for (int h = lane; h < p.hid; h += 32) {
    const float d = n1[h] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.hid + 1e-5f);
  const unsigned thr = (unsigned)(drop_p * 4294967296.0);
  const float inv = 1.0f / (1.0f - drop_p);
  for (int h = lane; h < p.hid; h += 32) {
    const float z = (n1[h] - mean) * rstd;
    float k1 = 1.0f, k2 = 1.0f;
    if (drop_p > 0.0f) {
      k1 = tw_keep(seed, blockIdx.y, row, h, 0, thr, inv);
      k2 = tw_keep(seed, blockIdx.y, row, h, 1, thr, inv);
    }
    n1[h] = (z * p.ln1w[h] + p.ln1b[h]) * k1;
    n2[h] = (z * p.ln2w[h] + p.ln2b[h]) * k2;
  }
  __syncwarp();
  for (int o = lane; o < p.out_dim; o += 32) {
    const float* w1 = p.W1 + (long long)o * p.hid;
    const float* w2 = p.W2 + (long long)o * p.hid;
    float m = p.b1[o], g = p.b2[o];
    for (int h = 0; h < p.hid; ++h) {
      m = fmaf(__ldg(w1 + h), n1[h], m);
      g = fmaf(__ldg(w2 + h), n2[h], g);
    }
    float sk;
    if (p.Ws) {
      const float* ws = p.Ws + (long long)o * p.in_dim;
      sk = p.bs[o];
      for (int i = 0; i < p.in_dim; ++i) sk = fmaf(__ldg(ws + i), x[i], sk);
    } else {
      sk = x[o];
    }
    Y[row * ldy + p.y_off + o] = m * sigmoidf_(g) + sk;
  }
}

// ---- backward (CTA = 32 rows x 8 warps) ------------------------------------------------------------------
__global__ void __launch_bounds__(TG_BWD_WARPS * 32) tower_group_bwd_kernel(const float* __restrict__ X, int ldx, int rows,
                                                                            const __grid_constant__ TowerGroup G, const float* __restrict__ A, int lda,
                                                                            const float* __restrict__ dY, int ldy, float* __restrict__ dA,
                                                                            float* __restrict__ dX, float drop_p, AcbSeed seed_s) {
  const unsigned long long seed = seed_s.get();
  extern __shared__ float sm[];
  const TowerT& p = G.t[blockIdx.y];
  const int in = p.in_dim, hid = p.hid, out = p.out_dim;
  float* sX = sm;                    // [R][in]
  float* sN1 = sX + TG_ROWS * in;    // [R][hid]  dropped affine-normalised activations of the main head
  float* sN2 = sN1 + TG_ROWS * hid;  // [R][hid]  ... of the gate head
  float* sDa = sN2 + TG_ROWS * hid;  // [R][hid]  d loss / d pre-GELU
  float* sDm = sDa + TG_ROWS * hid;  // [R][out]
  float* sDg = sDm + TG_ROWS * out;  // [R][out]
  float* sDy = sDg + TG_ROWS * out;  // [R][out]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long row_base = (long long)blockIdx.x * TG_ROWS;
  const int nrows = (int)min((long long)TG_ROWS, rows - row_base);
  const unsigned thr = (unsigned)(drop_p * 4294967296.0);
  const float inv = 1.0f / (1.0f - drop_p);

  float ln1w[TG_HPL], ln2w[TG_HPL], ln1b[TG_HPL], ln2b[TG_HPL];
  float g_ln1w[TG_HPL], g_ln1b[TG_HPL], g_ln2w[TG_HPL], g_ln2b[TG_HPL], g_b0[TG_HPL];
#pragma unroll
  for (int k = 0; k < TG_HPL; ++k) {
    const int h = lane + 32 * k;
    const bool ok = h < hid;
    ln1w[k] = ok ? p.ln1w[h] : 0.0f; ln1b[k] = ok ? p.ln1b[h] : 0.0f;
    ln2w[k] = ok ? p.ln2w[h] : 0.0f; ln2b[k] = ok ? p.ln2b[h] : 0.0f;
    g_ln1w[k] = g_ln1b[k] = g_ln2w[k] = g_ln2b[k] = g_b0[k] = 0.0f;
  }
  float g_b1 = 0.0f, g_b2 = 0.0f, g_bs = 0.0f;  // lane = output index (out <= 32)

  for (int r = wid; r < TG_ROWS; r += TG_BWD_WARPS) {
    float* n1 = sN1 + r * hid;
    float* n2 = sN2 + r * hid;
    float* da_s = sDa + r * hid;
    if (r >= nrows) {  // zero rows keep the tile products exact
      for (int i = lane; i < in; i += 32) sX[r * in + i] = 0.0f;
      for (int h = lane; h < hid; h += 32) { n1[h] = 0.0f; n2[h] = 0.0f; da_s[h] = 0.0f; }
      if (lane < out) { sDm[r * out + lane] = 0.0f; sDg[r * out + lane] = 0.0f; sDy[r * out + lane] = 0.0f; }
      continue;
    }
    const long long row = row_base + r;
    for (int i = lane; i < in; i += 32) sX[r * in + i] = X[row * ldx + (p.cols ? p.cols[i] : i)];
    // recompute the row from the saved pre-activation
    float a[TG_HPL], z[TG_HPL], k1[TG_HPL], k2[TG_HPL];
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < TG_HPL; ++k) {
      const int h = lane + 32 * k;
      a[k] = h < hid ? A[row * lda + p.a_off + h] : 0.0f;
      z[k] = h < hid ? gelu_erf(a[k]) : 0.0f;
      sum += z[k];
    }
    const float mean = warp_sum(sum) / (float)hid;
    float q = 0.0f;
#pragma unroll
    for (int k = 0; k < TG_HPL; ++k) {
      const int h = lane + 32 * k;
      z[k] = h < hid ? z[k] - mean : 0.0f;
      q += z[k] * z[k];
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)hid + 1e-5f);
#pragma unroll
    for (int k = 0; k < TG_HPL; ++k) {
      const int h = lane + 32 * k;
      z[k] *= rstd;
      k1[k] = k2[k] = 1.0f;
      if (drop_p > 0.0f && h < hid) {
        k1[k] = tw_keep(seed, blockIdx.y, row, h, 0, thr, inv);
        k2[k] = tw_keep(seed, blockIdx.y, row, h, 1, thr, inv);
      }
      if (h < hid) {
        n1[h] = (z[k] * ln1w[k] + ln1b[k]) * k1[k];
        n2[h] = (z[k] * ln2w[k] + ln2b[k]) * k2[k];
      }
    }
    __syncwarp();
    // heads (lane = output), their gradients
    float dm = 0.0f, dg = 0.0f, dyv = 0.0f;
    if (lane < out) {
      const float* w1 = p.W1 + (long long)lane * hid;
      const float* w2 = p.W2 + (long long)lane * hid;
      float m = p.b1[lane], g = p.b2[lane];
      for (int h = 0; h < hid; ++h) {
        m = fmaf(__ldg(w1 + h), n1[h], m);
        g = fmaf(__ldg(w2 + h), n2[h], g);
      }
      const float sg = sigmoidf_(g);
      dyv = dY[row * ldy + p.y_off + lane];
      dm = dyv * sg;
      dg = dyv * m * sg * (1.0f - sg);
      g_b1 += dm; g_b2 += dg; g_bs += dyv;
      sDm[r * out + lane] = dm; sDg[r * out + lane] = dg; sDy[r * out + lane] = dyv;
    }
    __syncwarp();
    // back through the two heads, the shared LayerNorm statistics and the GELU
    float dz[TG_HPL];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < TG_HPL; ++k) {
      const int h = lane + 32 * k;
      dz[k] = 0.0f;
      if (h < hid) {
        float d1 = 0.0f, d2 = 0.0f;
        for (int o = 0; o < out; ++o) {
          d1 = fmaf(__ldg(p.W1 + (long long)o * hid + h), sDm[r * out + o], d1);
          d2 = fmaf(__ldg(p.W2 + (long long)o * hid + h), sDg[r * out + o], d2);
        }
        d1 *= k1[k];
        d2 *= k2[k];
        g_ln1w[k] = fmaf(d1, z[k], g_ln1w[k]); g_ln1b[k] += d1;
        g_ln2w[k] = fmaf(d2, z[k], g_ln2w[k]); g_ln2b[k] += d2;
        dz[k] = d1 * ln1w[k] + d2 * ln2w[k];
        s1 += dz[k];
        s2 += dz[k] * z[k];
      }
    }
    s1 = warp_sum(s1) / (float)hid;
    s2 = warp_sum(s2) / (float)hid;
#pragma unroll
    for (int k = 0; k < TG_HPL; ++k) {
      const int h = lane + 32 * k;
      if (h < hid) {
        const float da = rstd * (dz[k] - s1 - z[k] * s2) * gelu_erf_grad(a[k]);
        da_s[h] = da;
        g_b0[k] += da;
        if (dA) dA[row * lda + p.a_off + h] = da;
      }
    }
    __syncwarp();
    if (dX) {  // gradient reaching the tower input (several towers may share columns: atomics)
      for (int i = lane; i < in; i += 32) {
        float d = 0.0f;
        if (p.Ws) {
          for (int o = 0; o < out; ++o) d = fmaf(__ldg(p.Ws + (long long)o * in + i), sDy[r * out + o], d);
        } else if (i < out) {
          d = sDy[r * out + i];
        }
        if (p.W0)
          for (int h = 0; h < hid; ++h) d = fmaf(__ldg(p.W0 + (long long)h * in + i), da_s[h], d);
        atomicAdd(dX + row * ldx + (p.cols ? p.cols[i] : i), d);
      }
    }
  }
  // vector gradients: registers -> global
#pragma unroll
  for (int k = 0; k < TG_HPL; ++k) {
    const int h = lane + 32 * k;
    if (h < hid) {
      atomicAdd(p.gln1w + h, g_ln1w[k]); atomicAdd(p.gln1b + h, g_ln1b[k]);
      atomicAdd(p.gln2w + h, g_ln2w[k]); atomicAdd(p.gln2b + h, g_ln2b[k]);
      if (p.W0) atomicAdd(p.gb0 + h, g_b0[k]);
    }
  }
  if (lane < out) {
    atomicAdd(p.gb1 + lane, g_b1);
    atomicAdd(p.gb2 + lane, g_b2);
    if (p.Ws) atomicAdd(p.gbs + lane, g_bs);
  }
  __syncthreads();
  // outer-product gradients of the 32-row tile
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < out * hid; idx += nt) {
    const int o = idx / hid, h = idx - o * hid;
    float a1 = 0.0f, a2 = 0.0f;
#pragma unroll 8
    for (int r = 0; r < TG_ROWS; ++r) {
      a1 = fmaf(sDm[r * out + o], sN1[r * hid + h], a1);
      a2 = fmaf(sDg[r * out + o], sN2[r * hid + h], a2);
    }
    atomicAdd(p.gW1 + idx, a1);
    atomicAdd(p.gW2 + idx, a2);
  }
  if (p.W0) {
    for (int idx = tid; idx < hid * in; idx += nt) {
      const int h = idx / in, i = idx - h * in;
      float acc = 0.0f;
#pragma unroll 8
      for (int r = 0; r < TG_ROWS; ++r) acc = fmaf(sDa[r * hid + h], sX[r * in + i], acc);
      atomicAdd(p.gW0 + idx, acc);
    }
  }
  if (p.Ws) {
    for (int idx = tid; idx < out * in; idx += nt) {
      const int o = idx / in, i = idx - o * in;
      float acc = 0.0f;
#pragma unroll 8
      for (int r = 0; r < TG_ROWS; ++r) acc = fmaf(sDy[r * out + o], sX[r * in + i], acc);
      atomicAdd(p.gWs + idx, acc);
    }
  }
}

int fill_group(TowerGroup& G, int n_towers, const long long* ptrs, const int* dims, bool with_grads, size_t* smem_out) {
  ACB_CHECK(n_towers >= 1 && n_towers <= TG_MAX_TOWERS && ptrs && dims, "acb_tower_group: %d towers (max %d)", n_towers, TG_MAX_TOWERS);
  G.n = n_towers;
  size_t smem = 0;
  for (int t = 0; t < n_towers; ++t) {
    const long long* q = ptrs + 25 * t;
    const int* d = dims + 5 * t;
    TowerT& T = G.t[t];
    T.cols = (const int*)q[0];
    T.in_dim = d[0]; T.hid = d[1]; T.out_dim = d[2]; T.y_off = d[3]; T.a_off = d[4]; T.pad_ = 0;
    const float** w = &T.W0;
    for (int j = 0; j < 12; ++j) w[j] = (const float*)q[1 + j];
    float** g = &T.gW0;
    for (int j = 0; j < 12; ++j) g[j] = (float*)q[13 + j];
    ACB_CHECK(T.in_dim > 0 && T.in_dim <= TG_MAX_IN && T.hid > 0 && T.hid <= TG_MAX_HID && T.out_dim > 0 && T.out_dim <= 32,
              "acb_tower_group: tower %d dims out of range (in=%d hid=%d out=%d; out <= 32)", t, T.in_dim, T.hid, T.out_dim);
    ACB_CHECK(T.ln1w && T.ln1b && T.W1 && T.b1 && T.ln2w && T.ln2b && T.W2 && T.b2, "acb_tower_group: tower %d: null parameter", t);
    ACB_CHECK((T.W0 == nullptr) == (T.b0 == nullptr) && (T.Ws == nullptr) == (T.bs == nullptr), "acb_tower_group: tower %d: weight without bias", t);
    ACB_CHECK(T.Ws != nullptr || T.in_dim == T.out_dim, "acb_tower_group: tower %d: identity skip needs in_dim == out_dim", t);
    if (with_grads) {
      ACB_CHECK(T.gln1w && T.gln1b && T.gW1 && T.gb1 && T.gln2w && T.gln2b && T.gW2 && T.gb2 && (!T.W0 || (T.gW0 && T.gb0)) && (!T.Ws || (T.gWs && T.gbs)),
                "acb_tower_group_bwd: tower %d: null gradient buffer", t);
    }
    const size_t need = (size_t)TG_ROWS * (T.in_dim + 3 * T.hid + 3 * T.out_dim) * sizeof(float);
    smem = need > smem ? need : smem;
  }
  if (smem_out) *smem_out = smem;
  return ACB_OK;
}

}  // namespace

extern "C" {

int acb_tower_group_fwd(const float* X, int ldx, int rows, int n_towers, const long long* ptrs, const int* dims, float* Y, int ldy, float* A,
                        int lda, float drop_p, long long seed, void* stream) {
  ACB_CHECK(X && Y && A && rows >= 0 && drop_p >= 0.0f && drop_p < 1.0f, "acb_tower_group_fwd: bad arguments");
  TowerGroup G;
  const int rc = fill_group(G, n_towers, ptrs, dims, false, nullptr);
  if (rc != ACB_OK) return rc;
  if (rows == 0) return ACB_OK;
  tower_group_fwd_kernel<<<dim3(cdiv(rows, 4), n_towers), 128, 0, (cudaStream_t)stream>>>(X, ldx, rows, G, Y, ldy, A, lda, drop_p, acb_seed(seed));
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_tower_group_bwd(const float* X, int ldx, int rows, int n_towers, const long long* ptrs, const int* dims, const float* A, int lda,
                        const float* dY, int ldy, float* dA, float* dX, float drop_p, long long seed, void* stream) {
  ACB_CHECK(X && A && dY && rows >= 0 && drop_p >= 0.0f && drop_p < 1.0f, "acb_tower_group_bwd: bad arguments");
  TowerGroup G;
  size_t smem = 0;
  const int rc = fill_group(G, n_towers, ptrs, dims, true, &smem);
  if (rc != ACB_OK) return rc;
  if (rows == 0) return ACB_OK;
  ACB_CHECK(smem <= 200 * 1024, "acb_tower_group_bwd: tile needs %zu bytes of shared memory", smem);
  if (dX) ACB_CUDA(cudaMemsetAsync(dX, 0, (size_t)rows * ldx * sizeof(float), (cudaStream_t)stream));
  ACB_CUDA(cudaFuncSetAttribute(tower_group_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tower_group_bwd_kernel<<<dim3(cdiv(rows, TG_ROWS), n_towers), TG_BWD_WARPS * 32, smem, (cudaStream_t)stream>>>(
      X, ldx, rows, G, A, lda, dY, ldy, dA, dX, drop_p, acb_seed(seed));
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
