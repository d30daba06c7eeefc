// Small-MLP kernels: gated residual metadata towers / MoE experts, top-2 mixture, late-fusion head.
// One warp per row, weights streamed through L1 (they are a few KB), fp32 FFMA, shuffle reductions.
#include "common.cuh"

namespace {

constexpr int TOWER_MAX_IN = 512;
constexpr int TOWER_MAX_HID = 256;
constexpr int TOWER_WARPS = 4;

struct TowerArgs {
  const float* X; int ldx; const int* cols; int in_dim, hid, out_dim;
  const float *W0, *b0, *ln1w, *ln1b, *W1, *b1, *ln2w, *ln2b, *W2, *b2, *Ws, *bs;
  float* Y; int ldy, y_off, rows;
  const float* S_pre; int lds, s_off;  // optional precomputed gelu(W0 x + b0) (experts: one GEMM for all start paths)
};

__global__ void __launch_bounds__(TOWER_WARPS * 32) tower_fwd_kernel(const TowerArgs p) {
  __shared__ float xs[TOWER_WARPS][TOWER_MAX_IN];
  __shared__ float ss[TOWER_WARPS][TOWER_MAX_HID];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * TOWER_WARPS + wid;
  if (row >= p.rows) return;
  float* x = xs[wid];
  float* s = ss[wid];
  for (int i = lane; i < p.in_dim; i += 32) x[i] = p.X[(long long)row * p.ldx + (p.cols ? p.cols[i] : i)];
  __syncwarp();
  // start_path: s = gelu(W0 x + b0)
  float sum = 0.0f;
  for (int h = lane; h < p.hid; h += 32) {
    float a;
    if (p.S_pre) {
      a = p.S_pre[(long long)row * p.lds + p.s_off + h];
    } else {
      const float* wr = p.W0 + (long long)h * p.in_dim;
      a = p.b0[h];
      for (int i = 0; i < p.in_dim; ++i) a = fmaf(__ldg(wr + i), x[i], a);
      a = gelu_erf(a);
    }
    s[h] = a;
    sum += a;
  }
  const float mean = warp_sum(sum) / (float)p.hid;
  float q = 0.0f;
  for (int h = lane; h < p.hid; h += 32) {
    const float d = s[h] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.hid + 1e-5f);
  __syncwarp();
  // normalised hidden (shared statistics, two affine sets) folded into the two output GEMVs.  Lanes run along the HIDDEN axis, so a
  // weight row is read as consecutive 128-byte lines (lanes along the outputs read it with a stride of `hid` floats: 32 sectors per
  // load, which made this kernel L1-throughput bound), and every output is a warp reduction; lane (o mod 32) keeps output o.
  constexpr int HPL = TOWER_MAX_HID / 32, OPL = 4;  // hidden values / outputs per lane
  float z1[HPL], z2[HPL];
#pragma unroll
  for (int k = 0; k < HPL; ++k) {
    const int h = lane + 32 * k;
    z1[k] = 0.0f;
    z2[k] = 0.0f;
    if (h < p.hid) {
      const float z = (s[h] - mean) * rstd;
      z1[k] = fmaf(z, p.ln1w[h], p.ln1b[h]);
      z2[k] = fmaf(z, p.ln2w[h], p.ln2b[h]);
    }
  }
  float mm[OPL], gg[OPL], sk[OPL];
#pragma unroll
  for (int j = 0; j < OPL; ++j) mm[j] = gg[j] = sk[j] = 0.0f;
  const bool wide_skip = p.Ws != nullptr && p.in_dim > 32;  // experts: 288 inputs -> lanes along the inputs as well
#pragma unroll 4  // (independent outputs: lets the loads of the next ones issue under the shuffle chains)
  for (int o = 0; o < p.out_dim; ++o) {
    const float* w1 = p.W1 + (long long)o * p.hid;
    const float* w2 = p.W2 + (long long)o * p.hid;
    float m = 0.0f, g = 0.0f, k3 = 0.0f;
#pragma unroll
    for (int k = 0; k < HPL; ++k) {
      const int h = lane + 32 * k;
      if (h < p.hid) {
        m = fmaf(__ldg(w1 + h), z1[k], m);
        g = fmaf(__ldg(w2 + h), z2[k], g);
      }
    }
    if (wide_skip) {
      const float* ws = p.Ws + (long long)o * p.in_dim;
      for (int i = lane; i < p.in_dim; i += 32) k3 = fmaf(__ldg(ws + i), x[i], k3);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      m += __shfl_xor_sync(0xffffffffu, m, off);
      g += __shfl_xor_sync(0xffffffffu, g, off);
      if (wide_skip) k3 += __shfl_xor_sync(0xffffffffu, k3, off);
    }
    if ((o & 31) == lane) {
#pragma unroll
      for (int j = 0; j < OPL; ++j)
        if (j == (o >> 5)) {
          mm[j] = m + p.b1[o];
          gg[j] = g + p.b2[o];
          sk[j] = k3 + (wide_skip ? p.bs[o] : 0.0f);
        }
    }
  }
#pragma unroll
  for (int j = 0; j < OPL; ++j) {
    const int o = lane + 32 * j;
    if (o < p.out_dim) {
      float skv = sk[j];
      if (!wide_skip) {
        if (p.Ws) {
          const float* ws = p.Ws + (long long)o * p.in_dim;
          skv = p.bs[o];
          for (int i = 0; i < p.in_dim; ++i) skv = fmaf(__ldg(ws + i), x[i], skv);
        } else {
          skv = x[o];
        }
      }
      p.Y[(long long)row * p.ldy + p.y_off + o] = mm[j] * sigmoidf_(gg[j]) + skv;
    }
  }
}

__global__ void moe_combine_kernel(const float* __restrict__ gate, const float* __restrict__ eo, float* __restrict__ out,
                                   int* __restrict__ top_idx, int B, int E, int C) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  // top-2 by value; ties resolved towards the lower index (torch.topk on distinct values is unambiguous)
  int i0 = 0;
  float v0 = gate[(long long)r * E];
  for (int e = 1; e < E; ++e) {
    const float v = gate[(long long)r * E + e];
    if (v > v0) { v0 = v; i0 = e; }
  }
  int i1 = -1;
  float v1 = -INFINITY;
  for (int e = 0; e < E; ++e) {
    if (e == i0) continue;
    const float v = gate[(long long)r * E + e];
    if (v > v1) { v1 = v; i1 = e; }
  }
  if (top_idx) { top_idx[2 * r] = i0; top_idx[2 * r + 1] = i1; }
  for (int c = 0; c < C; ++c) {
    float acc = 0.0f;
    for (int e = 0; e < E; ++e) {  // expert order e = 0..E-1 as in the reference loop
      if (e == i0) acc += v0 * eo[((long long)r * E + e) * C + c];
      else if (e == i1) acc += v1 * eo[((long long)r * E + e) * C + c];
    }
    out[(long long)r * C + c] = acc;
  }
}

struct FusionArgs {
  const float *p_in, *im_in, *s_in; int p_dim, im_dim, s_dim;
  const float *Wp, *bp, *Wim, *bim, *Ws, *bs, *Wfc, *bfc;
  int H, concat, num_classes; float* logits; float* emb_out; int B;
};

constexpr int FUSION_MAX_H = 256;

__device__ __forceinline__ void proj_norm(const float* in, int dim, const float* W, const float* b, int H, float* e, int lane) {
  float ss = 0.0f;
  if (dim >= 32) {
    // wide input (the 128-d photometry embedding): lanes along the input so that weight rows are read as full lines, one warp
    // reduction per output (lanes along the outputs read W with a stride of `dim` floats)
    float xr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xr[k] = lane + 32 * k < dim ? in[lane + 32 * k] : 0.0f;
#pragma unroll 4
    for (int h = 0; h < H; ++h) {
      const float* wr = W + (long long)h * dim;
      float a = 0.0f;
      if (dim <= 128) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (lane + 32 * k < dim) a = fmaf(__ldg(wr + lane + 32 * k), xr[k], a);
      } else {
        for (int i = lane; i < dim; i += 32) a = fmaf(__ldg(wr + i), in[i], a);
      }
      a = warp_sum(a) + b[h];  // (every lane holds the sum)
      if (lane == (h & 31)) {
        e[h] = a;
        ss += a * a;
      }
    }
  } else {
    for (int h = lane; h < H; h += 32) {
      const float* wr = W + (long long)h * dim;
      float a = b[h];
      for (int i = 0; i < dim; ++i) a = fmaf(__ldg(wr + i), in[i], a);
      e[h] = a;
      ss += a * a;
    }
  }
  const float inv = 1.0f / sqrtf(warp_sum(ss));
  __syncwarp();
  for (int h = lane; h < H; h += 32) e[h] *= inv;
  __syncwarp();
}

__global__ void __launch_bounds__(128) fusion_head_kernel(const FusionArgs p) {
  __shared__ float em[4][3][FUSION_MAX_H];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * 4 + wid;
  if (r >= p.B) return;
  float* ep = em[wid][0];
  float* ei = em[wid][1];
  float* es = em[wid][2];
  proj_norm(p.p_in + (long long)r * p.p_dim, p.p_dim, p.Wp, p.bp, p.H, ep, lane);
  proj_norm(p.im_in + (long long)r * p.im_dim, p.im_dim, p.Wim, p.bim, p.H, ei, lane);
  proj_norm(p.s_in + (long long)r * p.s_dim, p.s_dim, p.Ws, p.bs, p.H, es, lane);
  if (p.emb_out) {
    for (int h = lane; h < p.H; h += 32) {
      p.emb_out[((long long)0 * p.B + r) * p.H + h] = ep[h];
      p.emb_out[((long long)1 * p.B + r) * p.H + h] = ei[h];
      p.emb_out[((long long)2 * p.B + r) * p.H + h] = es[h];
    }
  }
  const int F = p.concat ? 3 * p.H : p.H;
  for (int c = 0; c < p.num_classes; ++c) {
    const float* wr = p.Wfc + (long long)c * F;
    float a = 0.0f;
    for (int h = lane; h < p.H; h += 32) {
      if (p.concat) {
        a = fmaf(wr[h], ep[h], a);
        a = fmaf(wr[p.H + h], ei[h], a);
        a = fmaf(wr[2 * p.H + h], es[h], a);
      } else {
        a = fmaf(wr[h], (ep[h] + ei[h] + es[h]) / 3.0f, a);
      }
    }
    a = warp_sum(a);
    if (lane == 0) p.logits[(long long)r * p.num_classes + c] = a + p.bfc[c];
  }
}

}  // namespace

extern "C" {

int acb_tower_fwd(const float* X, int ldx, const int* cols, int in_dim, int hid, int out_dim, const float* W0,
                  const float* b0, const float* ln1w, const float* ln1b, const float* W1, const float* b1,
                  const float* ln2w, const float* ln2b, const float* W2, const float* b2, const float* Ws,
                  const float* bs, float* Y, int ldy, int y_off, int rows, const float* S_pre, int lds, int s_off, void* stream) {
  ACB_CHECK(X && Y && (S_pre || (W0 && b0)) && ln1w && ln1b && W1 && b1 && ln2w && ln2b && W2 && b2, "acb_tower_fwd: null argument");
  ACB_CHECK(in_dim > 0 && in_dim <= TOWER_MAX_IN && hid > 0 && hid <= TOWER_MAX_HID && out_dim > 0 && out_dim <= 128,
            "acb_tower_fwd: dims out of range (in=%d hid=%d out=%d)", in_dim, hid, out_dim);
  ACB_CHECK(Ws != nullptr || in_dim == out_dim, "acb_tower_fwd: identity skip needs in_dim == out_dim");
  if (rows == 0) return ACB_OK;
  TowerArgs p{X, ldx, cols, in_dim, hid, out_dim, W0, b0, ln1w, ln1b, W1, b1, ln2w, ln2b, W2, b2, Ws, bs, Y, ldy, y_off, rows, S_pre, lds, s_off};
  tower_fwd_kernel<<<cdiv(rows, TOWER_WARPS), TOWER_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_moe_combine(const float* gate, const float* expert_out, float* out, int* top_idx, int B, int E, int C, void* stream) {
  ACB_CHECK(gate && expert_out && out && B > 0 && E >= 2 && C > 0, "acb_moe_combine: bad arguments");
  moe_combine_kernel<<<cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(gate, expert_out, out, top_idx, B, E, C);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_fusion_head(const float* p_in, int p_dim, const float* im_in, int im_dim, const float* s_in, int s_dim,
                    const float* Wp, const float* bp, const float* Wim, const float* bim, const float* Ws,
                    const float* bs, const float* Wfc, const float* bfc, int H, int concat, int num_classes,
                    float* logits, float* emb_out, int B, void* stream) {
  ACB_CHECK(p_in && im_in && s_in && Wp && bp && Wim && bim && Ws && bs && Wfc && bfc && logits, "acb_fusion_head: null argument");
  ACB_CHECK(H > 0 && H <= FUSION_MAX_H && B > 0 && num_classes > 0, "acb_fusion_head: bad dims");
  FusionArgs p{p_in, im_in, s_in, p_dim, im_dim, s_dim, Wp, bp, Wim, bim, Ws, bs, Wfc, bfc, H, concat, num_classes, logits, emb_out, B};
  fusion_head_kernel<<<cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(p);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
