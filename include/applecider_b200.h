/* applecider_b200 — C-ABI of the B200 (sm_100a) hot path of the AppleCiDEr multimodal classifier.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The host layer
 * (applecider_b200/*.py, torch.nn.Modules with the reference's forward signatures and
 * state_dict keys) binds these symbols with ctypes and passes tensor.data_ptr() values and the
 * current CUDA stream; INTEGRATION.md shows the same stub added to the reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - activations are channels-last ([rows, C] row-major); weights keep PyTorch layouts
 *     ((out, in[, k])) unless a "packed" layout is stated;
 *   - dtype tags: ACB_F32 / ACB_BF16; bf16 tensors are raw uint16 storage of __nv_bfloat16;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host;
 *   - return value 0 = success, <0 = error (message: acb_last_error()); there is NO CPU
 *     fallback anywhere behind this interface.
 *
 * Each entry point cites the reference code (paths relative to the reference root) it replaces.
 */
#ifndef APPLECIDER_B200_H
#define APPLECIDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACB_F32 0
#define ACB_BF16 1

#define ACB_ACT_NONE 0
#define ACB_ACT_RELU 1
#define ACB_ACT_GELU 2 /* exact erf GELU (torch F.gelu default) */
#define ACB_ACT_TANH 3
#define ACB_ACT_SIGMOID 4
#define ACB_ACT_SOFTPLUS 5 /* log(1 + exp(x)), linear above 20 like torch (legacy redshift head, SpectraNetRedshift.py:112) */

#define ACB_RES_NONE 0
#define ACB_RES_ADD 1 /* C = res + gamma[n] * v   (gamma may be NULL = 1) */
#define ACB_RES_MUL 2 /* C = res * v */
#define ACB_RES_MUL_GELU_GRAD 3 /* C = v * gelu'(res): gradient GEMM fused with the backward of h = gelu(res) (bf16 path) */

#define ACB_OK 0
#define ACB_ERR_INVALID (-1)
#define ACB_ERR_CUDA (-2)
#define ACB_ERR_UNSUPPORTED (-3)

/* ---- library management -------------------------------------------------------------------- */
const char* acb_last_error(void);
int acb_version(void);
long long acb_launch_count(void); /* kernels launched by this library since the last reset */
void acb_reset_launch_count(void);
/* Device-resident epoch of the stochastic kernels (dropout, attention dropout, Time2Vec dropout, MPT masking): when set
 * (NULL clears it), every such kernel uses seed + f(*dev_ptr) instead of the bare host seed.  A captured CUDA graph bakes
 * host seeds in; incrementing *dev_ptr inside the graph gives each replay fresh masks while the forward and backward
 * kernels of one replay still regenerate the same ones.  Dropout sites: nn.TransformerEncoderLayer (HyraxBaselineCLS.py:26-33),
 * astrominn.py:48-54,268, spectranet.py:148. */
int acb_set_seed_epoch_ptr(const unsigned long long* dev_ptr);

/* ---- GEMM / implicit-GEMM conv1d ---------------------------------------------------------------
 * v[m,n] = sum_k A(m,k) * Bw[n*ldb + k] (+ bias[n]); C = epilogue(act(v)).
 * Plain GEMM: conv_L = 0, A is [M, lda].  Implicit conv1d ("same" zero padding): conv_L = L > 0,
 * A is the channels-last signal [M/L, L, conv_Cin]; k = tap*conv_Cin + ci reads row l+tap-conv_pad.
 * Replaces nn.Linear / nn.Conv1d / nn.Conv2d-as-patch-GEMM call sites:
 *   src/applecider/models/HyraxBaselineCLS.py:60,78 ; spectranet.py:29,37 ; astrominn.py:12-41,270-295.
 */

/* fp32 CUDA-core path (parity mode, and every tiny-N head). */
int acb_gemm_f32(const float* A, const float* Bw, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                 int conv_L, int conv_Cin, int conv_pad, const float* bias, int act, const float* res, int ldr,
                 const float* gamma, int res_mode, void* stream);

/* bf16 tcgen05 path: TMA-fed UMMA (M=128 x BN tiles, fp32 accumulators in TMEM).
 * A is viewed as [nbatch, L, Cin] with element strides (a_batch_stride, a_row_stride, 1) — strides may
 * overlap (polyphase view of a 1-channel signal); K = taps*Cin with tap shifting the row by tap-pad and
 * out-of-range rows reading zeros (TMA OOB fill).  Plain GEMM: nbatch=1, L=M, Cin=K, taps=1, pad=0.
 * Bw is [N, ldb] bf16, K-major.  tile_kb_host (optional, host memory, [N/BN][2]) restricts every
 * N tile to a range of 64-wide K blocks (zero weights outside are never touched).
 * colblk_off_host (optional, host memory, [N/64]) remaps each 64-column block of the result to
 * column offset colblk_off[j] of C (default j*64).  pool4 != 0 max-pools groups of 4 consecutive rows.
 * m_valid_dev (optional) holds the number of valid rows on the device (tiles beyond it exit).
 * pre_out (optional, bf16, layout of C): second output = accumulator + bias BEFORE the activation / residual (training:
 * the tensor the backward of the activation needs, written by the same epilogue instead of a separate kernel). */
int acb_gemm_bf16(const void* A, const void* Bw, void* C, int c_dtype, int nbatch, int L, int Cin, int taps, int pad,
                  long long a_batch_stride, long long a_row_stride, int N, int ldb, int ldc, int bn,
                  const int* tile_kb_host, const int* colblk_off_host, const float* bias, int act, const void* res,
                  int res_dtype, int ldr, const float* gamma, int res_mode, int pool4, const int* m_valid_dev,
                  void* pre_out, void* stream);
/* SpectraNet block tail without the normalised activation ever reaching HBM (spectranet.py:36-40: norm -> GELU -> downsample
 * -> MaxPool1d(4)):
 *  - acb_gemm_bf16_stats: the conv GEMM above that ALSO writes, per output row and per N tile (ceil(N/bn) tiles), the sum and
 *    the sum of squares of that tile's columns: row_stats[rows][tiles][2] fp32 -- the LayerNorm statistics, for free, from
 *    the fp32 accumulators in the epilogue.
 *  - acb_gemm_ln_bf16: C = [maxpool4 over row quadruples]( gelu(LayerNorm_K(A)) Bw^T + bias ).  A[M,K] is the stored conv
 *    output; each TMA-loaded 128 x 64 tile is normalised in place in shared memory (mean / rstd from row_stats, weight /
 *    bias ln_w / ln_b over the K columns, erf-GELU via the bf16-accurate tanh form) before the tcgen05 MMA reads it.
 *    K % 64 == 0, N % 128 == 0. */
int acb_gemm_bf16_stats(const void* A, const void* Bw, void* C, int c_dtype, int nbatch, int L, int Cin, int taps, int pad,
                        long long a_batch_stride, long long a_row_stride, int N, int ldb, int ldc, int bn,
                        const int* tile_kb_host, const float* bias, float* row_stats, void* stream);
int acb_gemm_ln_bf16(const void* A, const void* Bw, void* C, int c_dtype, long long M, int K, int N, int ldc, const float* bias,
                     int pool4, const float* row_stats, int parts, const float* ln_w, const float* ln_b, float ln_eps, void* stream);

/* SpectraNet block front half fused on tcgen05 (spectranet.py:29-35): three same-padded Conv1d (implicit GEMM, packed
 * weights Bw [b_rows, ldb], per-conv K-block ranges kb_ranges_host[3][2]) + bias + LayerNorm over the concatenated
 * channels of every position + GELU -> bf16 out[out_rows, 3*(128/ng)].  Sub-tile j (= conv j, 128 accumulator columns)
 * reads weight rows brow_base_host[j] + blockIdx.y*brow_stride_y; ng = LayerNorm groups per GEMM row (stage 0 polyphase:
 * 2 phases per CTA, output row = m*row_mul + blockIdx.y*row_add_y + g).  A geometry as in acb_gemm_bf16. 
 * down_w / down_bias / down_out (optional, all or none; stage-0 polyphase form with >= 148 signal windows only): also fuse the
 * block's 1x1 downsample Conv1d(3*64 -> 64) (down_w bf16 [64,192], down_bias f32 [64]) and the first half of MaxPool1d(4):
 * down_out[out_rows/2, 64] bf16 = max over positions (2i, 2i+1) of conv1x1(gelu(LN(...))) + bias; `out` is then not written
 * (may be NULL) -- the [out_rows, 192] activation never reaches HBM (spectranet.py:36-40). */
int acb_spectra_conv_ln_bf16(const void* A, const void* Bw, void* out, int nbatch, int L, int Cin, int taps, int pad,
                             long long a_batch_stride, long long a_row_stride, int ldb, int b_rows, const int* kb_ranges_host,
                             const int* brow_base_host, int brow_stride_y, int grid_y, int ng, long long row_mul,
                             long long row_add_y, long long out_rows, const float* bias, const float* gamma,
                             const float* beta, float eps, const void* down_w, const float* down_bias, void* down_out, void* stream);

/* Weight packing (derived, non-persistent buffers; refreshed after load_state_dict / optimizer step).
 * out[co*row_stride + (tap + tap_off)*Cin + ci] = w[co, ci, tap]     (w = PyTorch (Cout, Cin, k)) */
int acb_pack_conv_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int k, long long row_stride,
                         int tap_off, void* stream);
/* polyphase packing of a 1-input-channel conv (spectranet.py stage 0): rows (r, co), columns kk:
 * out[(row_off + r*rows_per_phase + co)*row_stride + kk] = w[co,0,kk - r + pad - halo], 0 elsewhere */
int acb_pack_polyphase_weight(const float* w, void* out, int out_dtype, int Cout, int k, int phases, int halo,
                              int rows_per_phase, int row_off, long long row_stride, int Kp, void* stream);
int acb_cast(const void* in, int in_dtype, void* out, int out_dtype, long long n, void* stream);
/* out[c*R + r] = in[r*C + c] */
int acb_transpose(const void* in, int in_dtype, void* out, int out_dtype, int R, int C, void* stream);
/* the fp32 -> bf16 transposes of MANY weight matrices in one launch (the W^T operands of every nn.Linear's input-gradient GEMM,
 * refreshed together after an optimizer step): device arrays src_ptrs[n] (const float*), dst_ptrs[n] (bf16*), meta[n][4] =
 * {R, C, ceil(C / 32), first global 32 x 32 tile of the job}; total_tiles = sum over jobs of ceil(R / 32) * ceil(C / 32). */
int acb_transpose_batch(const long long* src_ptrs, const long long* dst_ptrs, const int* meta, int n_jobs, int total_tiles, void* stream);
/* out[b*out_stride + lead + i] = in[b*L + i] (bf16/f32), buffer pre-zeroed by the caller */
int acb_pad_signal(const float* in, void* out, int out_dtype, int nb, int L, long long out_stride, int lead,
                   void* stream);

/* ---- row-wise LayerNorm (+ fused pre-GELU / residual / post-activation) --------------------------
 * v = x[r,:]; if pre_gelu v = gelu(v); if res v += res[r,:]; y = (v-mean)*rstd*w + b; y = act(y).
 * nn.LayerNorm call sites: HyraxBaselineCLS.py:26-33,79 ; spectranet.py:32,145 ; astrominn.py:23-34,48-54 ;
 * timm LayerNorm/LayerNorm2d (eps 1e-6). */
int acb_layernorm(const void* x, int x_dtype, const void* res, int res_dtype, const float* w, const float* b,
                  void* y, int y_dtype, long long rows, int C, float eps, int pre_gelu, int post_act, void* stream);
/* the same with the row count read on the device: rows_dev (optional) caps `rows`, so a capacity-sized token matrix
 * (acb_photo_compact) costs no work for the rows past the packed count and nothing is read back to the host. */
int acb_layernorm_n(const void* x, int x_dtype, const void* res, int res_dtype, const float* w, const float* b,
                    void* y, int y_dtype, long long rows, int C, float eps, int pre_gelu, int post_act, const int* rows_dev,
                    void* stream);

/* ---- photometry transformer pieces --------------------------------------------------------------- */
/* pad[B,L] (bool bytes, nonzero = padding) -> cu_seqlens[B+1] (tokens incl. CLS) and src_idx[T]
 * (source row b*L+l of every packed token, -1-b for the CLS token of sequence b).  src_idx has the full capacity of
 * B*(L+1) entries; those past cu_seqlens[B] are "dead" (0x80808080): acb_photo_embed writes zero rows for them, so a
 * caller may size the token matrix by any upper bound of the packed count without reading it back from the device.
 * capacity (0 = B*(L+1)) is that bound: cu_seqlens is clamped to it, so a too-small bound truncates instead of overrunning.
 * HyraxBaselineCLS.py:71-78 (CLS prepend + key-padding mask); equals PyTorch's nested-tensor packing. */
int acb_photo_compact(const uint8_t* pad, int B, int L, int capacity, int* cu_seqlens, int* src_idx, void* stream);
/* packed tokens h[t,:] = in_proj(x[src]) + Time2Vec(x[src,0]) or cls_tok  (HyraxBaselineCLS.py:60-72,
 * Time2Vec.py:63-72).  D = d_model. */
int acb_photo_embed(const float* x, const int* src_idx, const int* total_dev, int max_tokens, int D, const float* w_in,
                    const float* b_in, const float* w0, const float* b0, const float* w, const float* b,
                    const float* cls_tok, float te_drop_p, long long te_seed, void* out, int out_dtype, void* stream);
/* fused masked varlen multi-head attention over packed tokens; qkv[T,3D] rows = [q|k|v], head h uses
 * columns h*dh..; softmax(q k^T / sqrt(dh)) v with fp32 math (drop_p > 0: training-time dropout on the
 * probabilities, mask = counter hash of (seed, sequence, head, i, j), regenerated in the backward).  nn.MultiheadAttention inside
 * nn.TransformerEncoderLayer (HyraxBaselineCLS.py:26-33,78). dh must be 16. */
int acb_attention_varlen(const void* qkv, int dtype, const int* cu_seqlens, int B, int n_heads, int dh,
                         int max_seqlen, float drop_p, long long seed, void* out, void* stream);
/* the same attention on tcgen05 for bf16 tokens: S = QK^T is one UMMA per (head, 128-query tile) into TMEM, fp32 softmax
 * per accumulator row, P V accumulated in TMEM; operands are no-swizzle K-major core matrices written by the CTA. */
int acb_attention_varlen_tc(const void* qkv, const int* cu_seqlens, int B, int n_heads, int dh, int max_seqlen, float drop_p,
                            long long seed, void* out, void* stream);
/* Packed multi-sequence attention (the default bf16 path): same arithmetic as acb_attention_varlen_tc, but
 *  - acb_attention_plan (once per batch, reused by every layer and by the backward) packs consecutive whole sequences into
 *    tiles of <= 128 token rows: plan[0] = tiles, plan[1] = long sequences (> 128 tokens), plan[2+2t], plan[3+2t] = first
 *    sequence / sequence count of tile t, plan[2+2*max_tiles+i] = i-th long sequence.  plan holds 2 + 2*max_tiles + B ints;
 *    max_tiles >= min(B, 2*(total_rows/128) + total_rows/129 + 2) is a host-side bound, nothing is read back;
 *  - acb_attention_packed: one CTA per (tile, 4 heads); Q/K/V of the tile arrive by TMA in the canonical UMMA layout, S and
 *    the per-head O live in TMEM, the block-diagonal mask is each row's own key range; long sequences run the
 *    per-(sequence, head) kernel over the plan's list.  total_rows = rows of the qkv / out matrices (capacity).
 * nn.MultiheadAttention inside nn.TransformerEncoderLayer with src_key_padding_mask (HyraxBaselineCLS.py:26-33,78). */
int acb_attention_plan(const int* cu_seqlens, int B, int max_tiles, int* plan, void* stream);
int acb_attention_packed(const void* qkv, const int* cu_seqlens, const int* plan, int B, int max_tiles, long long total_rows,
                         int n_heads, int dh, int max_seqlen, float drop_p, long long seed, void* out, void* stream);
/* backward of acb_attention_packed on tcgen05 (bf16 qkv / dout / dqkv): S and dP = dO V^T recomputed into TMEM, softmax and
 * dS = P (dP - sum P dP) / sqrt(dh) in registers, dV = P^T dO, dQ = dS K, dK = dS^T Q by N = 16 UMMAs (P^T / dS^T are the
 * shared-memory P / dS tiles read MN-major); the same plan and dropout hash as the forward; long sequences use the
 * per-(sequence, head) CUDA-core backward over the plan's list.  dqkv rows outside every sequence are left untouched. */
int acb_attention_packed_bwd(const void* qkv, const void* dout, const int* cu_seqlens, const int* plan, int B, int max_tiles,
                             long long total_rows, int n_heads, int dh, int max_seqlen, float drop_p, long long seed, void* dqkv,
                             void* stream);
/* out[b,:] = x[cu_seqlens[b],:]  (CLS read-out z[:,0], HyraxBaselineCLS.py:79) */
int acb_gather_cls(const void* x, int dtype, const int* cu_seqlens, int B, int D, float* out, void* stream);

/* ---- ConvNeXt-T pieces (timm convnext_tiny via astrominn.py:12-17; channels-last activations) ---- */
/* Fused MLP tail of a ConvNeXt block (inference, bf16):  out = res + gamma * (fc2(gelu(fc1(y) + b1)) + b2)  with y[M,C] the
 * LayerNorm-ed depthwise-conv output, w1[4C,C], w2[C,4C] bf16 row-major, res/out[M,C] bf16.  The [M,4C] hidden activation
 * stays in shared memory / TMEM (one 64- or 128-column chunk at a time).  C = 96 or 192 (stages 0 and 1, 98 % of the rows);
 * other widths return ACB_ERR_UNSUPPORTED and the caller keeps the two-GEMM path. */
int acb_convnext_mlp_bf16(const void* y, const void* res, const void* w1, const float* b1, const void* w2, const float* b2,
                          const float* gamma, void* out, long long M, int C, void* stream);
/* The same kernel as the transformer feed-forward block of nn.TransformerEncoderLayer (HyraxBaselineCLS.py:26-33):
 * out = x + linear2(relu(linear1(x) + b1)) + b2, d_model C = 128, width 4C; `ones` = C ones; rows_dev (optional) = device row
 * count of a capacity-sized token matrix (128-row tiles past it exit). */
int acb_ffn_relu_bf16(const void* x, const void* w1, const float* b1, const void* w2, const float* b2, const float* ones, void* out,
                      long long M, int C, const int* rows_dev, void* stream);
/* NCHW f32 image -> patch matrix [B*Ho*Wo, Cin*p*p] (k = ci*p*p + ky*p + kx), Ho = H/p (floor). */
int acb_patchify_nchw(const float* img, int B, int Cin, int H, int W, int p, void* out, int out_dtype, void* stream);
/* depthwise 7x7 (pad 3) + LayerNorm over C (eps) : x,y = [B,H,W,C]; w = (C,1,7,7), b = (C). */
int acb_dwconv7_ln(const void* x, int dtype, const float* w, const float* b, const float* ln_w, const float* ln_b,
                   float eps, void* y, int B, int H, int W, int C, void* stream);
/* LayerNorm2d(eps) then 2x2/stride-2 patch gather: x=[B,H,W,C] -> out=[B*(H/2)*(W/2), 4*C],
 * column (ky*2+kx)*C + c  (weights must be packed in the same (ky,kx,ci) order). */
int acb_ln_patch2(const void* x, int dtype, const float* ln_w, const float* ln_b, float eps, void* out, int B, int H,
                  int W, int C, void* stream);
/* global average pool over HW then LayerNorm(C, eps): x=[B,HW,C] -> out[B,C] f32 (timm head). */
int acb_gap_ln(const void* x, int dtype, const float* ln_w, const float* ln_b, float eps, float* out, int B, int HW,
               int C, void* stream);
/* out[b, (ky*2+kx)... generic weight re-layout (Cout,Cin,kh,kw) -> [Cout, kh*kw*Cin] ((ky,kx,ci) order) */
int acb_pack_conv2d_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int kh, int kw, void* stream);

/* ---- SpectraNet pooling (spectranet.py:25,38-40,163) ------------------------------------------ */
int acb_maxpool4_cl(const void* x, int dtype, void* y, int B, int L, int C, void* stream); /* -> [B, L/4, C] */
int acb_globalmax_cl(const void* x, int dtype, float* y, int B, int L, int C, void* stream); /* -> [B, C] f32 */
/* y[r,:] = max(x[2r,:], x[2r+1,:]) for bf16 rows of C (C % 8 == 0) channels: finishes MaxPool1d(4) after the fused stage-0
 * kernel (acb_spectra_conv_ln_bf16 with down_w) has max-ed the two positions each CTA owns. */
int acb_pairmax_bf16(const void* x, void* y, long long rows_out, int C, void* stream);
/* legacy spectra encoder ("variant B", _archive/notebooks/brew_cider.py:611-636): x[B,L,C] f32 channels-last ->
 * y[B, L/4, 3C] = [max | mean | min] over windows of 4 positions (cat of MaxPool1d, AvgPool1d, -MaxPool1d(-x)). */
int acb_tripool4_cl(const float* x, float* y, int B, int L, int C, void* stream);
/* the same for fp32 or bf16 tensors, and its backward: dx[l] = dmax*[l is the first argmax] + davg/4 + dmin*[l is the first
 * argmin] (torch's MaxPool1d routing); positions past 4*(L/4) get zero. */
int acb_tripool4(const void* x, int x_dtype, void* y, int y_dtype, int B, int L, int C, void* stream);
int acb_tripool4_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype, int B, int L, int C,
                     void* stream);
/* ---- BatchNorm1d over channels-last [rows, C] (legacy variant-B stages, brew_cider.py:601-604) ------------------------
 * acb_bn_stats      sums[0:C] = sum_r y[r,c], sums[C:2C] = sum_r y[r,c]^2 (fp32).
 * acb_bn_finalize   training: mean / biased var from sums -> scale = w*rstd, shift = b - mean*scale (and mean, rstd), and
 *                   running_mean / running_var blended with `momentum` (unbiased var), as nn.BatchNorm1d.train();
 *                   eval (training = 0): the same from the running statistics (sums unused).
 * acb_affine_res_act      out = act(res + scale[c]*y + shift[c])   (BN + skip projection + GELU of the block in one pass)
 * acb_affine_res_act_bwd  dpre = dout * act'(pre) with pre recomputed; sums[0:C] = sum dpre, sums[C:2C] = sum dpre*xhat
 *                         (= d bias, d weight of the BatchNorm); dpre is also the gradient of `res`.
 * acb_bn_bwd_apply  dy = scale * (dpre - sums[c]/rows - xhat*sums[C+c]/rows) in training, scale*dpre in eval. */
int acb_bn_stats(const void* y, int y_dtype, long long rows, int C, float* sums, void* stream);
int acb_bn_finalize(const float* sums, long long rows, int C, const float* w, const float* b, float eps, float momentum, int training,
                    float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* rstd, void* stream);
int acb_affine_res_act(const void* y, int y_dtype, const float* scale, const float* shift, const void* res, int res_dtype, int act,
                       void* out, int out_dtype, long long rows, int C, void* stream);
int acb_affine_res_act_bwd(const void* y, int y_dtype, const float* scale, const float* shift, const void* res, int res_dtype, int act,
                           const void* dout, int dout_dtype, const float* mean, const float* rstd, void* dpre, int dpre_dtype,
                           float* sums, long long rows, int C, void* stream);
int acb_bn_bwd_apply(const void* y, int y_dtype, const void* dpre, int dpre_dtype, const float* scale, const float* mean,
                     const float* rstd, const float* sums, int training, void* dy, int dy_dtype, long long rows, int C, void* stream);

/* ---- metadata towers / MoE / fusion head -------------------------------------------------------- */
/* ResidualTowerBlock (astrominn.py:44-64), eval: s = gelu(W0 x + b0); y = (W1 ln1(s) + b1) * sigmoid(W2 ln2(s) + b2)
 * + (Ws x + bs | x).  x = X[r, cols[i]] (cols == NULL: identity), Y[r, y_off + o].  hid <= 256, in <= 512, out <= 32.
 * S_pre (optional): precomputed s = gelu(W0 x + b0) at S_pre[r*lds + s_off + h] (one GEMM for all experts). */
int acb_tower_fwd(const float* X, int ldx, const int* cols, int in_dim, int hid, int out_dim, const float* W0,
                  const float* b0, const float* ln1w, const float* ln1b, const float* W1, const float* b1,
                  const float* ln2w, const float* ln2b, const float* W2, const float* b2, const float* Ws,
                  const float* bs, float* Y, int ldy, int y_off, int rows, const float* S_pre, int lds, int s_off, void* stream);
/* top-2-of-E sigmoid-gated mixture (astrominn.py:270-295): gate[B,E], expert_out[E][B,C] (stride E*C per row:
 * expert_out[r*E*C + e*C + c]) -> out[B,C]; also writes the selected indices (top_idx[B,2]) for parity checks. */
int acb_moe_combine(const float* gate, const float* expert_out, float* out, int* top_idx, int B, int E, int C,
                    void* stream);
/* late fusion (brew_cider.py:834-860): three Linear -> L2 normalise -> avg|concat -> fc.
 * emb_out (optional) = [3][B,H] (p, im, s). */
int acb_fusion_head(const float* p_in, int p_dim, const float* im_in, int im_dim, const float* s_in, int s_dim,
                    const float* Wp, const float* bp, const float* Wim, const float* bim, const float* Ws,
                    const float* bs, const float* Wfc, const float* bfc, int H, int concat, int num_classes,
                    float* logits, float* emb_out, int B, void* stream);
int acb_softmax_rows(const float* x, float* y, int rows, int C, void* stream);

/* ---- array-level preprocessing ------------------------------------------------ */
/* P1  datasets/photo_dataset.py:85-101,117-152 + models/HyraxBaselineCLS.py:157.
 * raw = concatenated (L_i,5) f32 rows [dt, dt_prev, band, logflux, logflux_err]; offsets[B+1] (int64).
 * horizon cut dt <= horizon -> log1p(dt), log1p(dt_prev), logf, logfe, one-hot band -> pad|truncate to
 * max_len -> (x - mean)/(std + 1e-8) on channels 0..3 (padding rows too).  x[B,max_len,7] f32,
 * mask[B,max_len] (1 = padding), lengths[B] (optional). */
int acb_prep_lightcurve(const float* raw, const long long* offsets, int B, float horizon, const float* mean,
                        const float* stdv, int max_len, float* x, uint8_t* mask, int* lengths, void* stream);
/* P2  preprocessing_utils/preprocess_multimodal.py:84-111,176-180,291-336.
 * Per object (detections time-ordered, offsets[B+1]): mag->flux, per-band greedy window merge (window
 * dt_days from the anchor, weights 1/(err+1e-8), fp64), 3-way merge by time, event features
 * dt, dt_prev, band_id (0 g,1 r,2 i), log10(clip(flux,1e-6)), flux_err/(ln10*flux).  Outputs use the input
 * offsets (an object has at most as many events as detections); n_events[B] = events per object.
 * tmp = 3*total doubles, tmp_b = total bytes of scratch. */
int acb_prep_events(const double* mjd, const double* mag, const double* magerr, const int* fid, const long long* offsets,
                    int B, long long total, double dt_days, double* tmp, signed char* tmp_b, float* dt, float* dt_prev,
                    signed char* band_id, float* logflux, float* logflux_err, int* n_events, void* stream);
/* P3  preprocess_multimodal.py:146-170 (_interp_with_extrap), :135-143 (_mad), :598-609.
 * Ragged spectra (wavelength, flux f64; any order; non-finite samples dropped) -> linear interpolation with
 * linear extrapolation onto grid[n_grid] (f32 wavelengths) -> subtract mean, divide by MAD (fallback std, 1)
 * -> out[B,n_grid] f32.  max_n = longest input spectrum. */
int acb_prep_spectrum_resample(const double* wl, const double* fx, const long long* offsets, int B, int max_n,
                               const float* grid, int n_grid, float* out, void* stream);
/* the same, also returning the searchsorted(side='left') position of every grid point among the spectrum's finite, sorted
 * wavelengths (idx_out[B, n_grid], -1 for spectra with < 2 finite samples): the integer part of the interpolation, which must
 * equal numpy's bit for bit (scipy interp1d picks the interval [idx-1, idx] clipped to [1, n-1], preprocess_multimodal.py:146-170). */
int acb_prep_spectrum_resample_idx(const double* wl, const double* fx, const long long* offsets, int B, int max_n,
                                   const float* grid, int n_grid, float* out, int* idx_out, void* stream);
/* P4  datasets/image_and_metadata_dataset.py:78-99; Fusion_Dataset.ipynb cell 0.
 * img[B,C,H,W] f32 -> centre crop [i1:i2] (i1 = int((H-cutout_size)/2), i2 = H-i1) -> mode 0: per-channel
 * lower-median subtraction and division by (unbiased std + 1e-8); mode 1: division by the L2 norm over all
 * channels; mode 2: notebook variant (true median, population std, std<=1e-8 -> 1).  out[B,C,S,S]. */
int acb_prep_cutout_norm(const float* img, int B, int C, int H, int W, int cutout_size, int mode, float* out, void* stream);
/* P5  preprocess_multimodal.py:863-895.  Column mean and population std (clipped at 0) of data[rows,F] f32
 * from streamed sums / sums of squares (fp64 accumulation).  work = 2*F doubles of scratch. */
int acb_feature_stats(const float* data, long long rows, int F, double* work, float* mean, float* stdv, void* stream);

/* ---- backward / training kernels (gradients of the ops above; torch autograd is only the tape) ---------- */
/* C[m,n] (+)= sum_k A(m,k) B(n,k) with arbitrary element strides (dgrad: B = W read column-wise; wgrad: A = dY,
 * B = X read column-wise).  convT: A is the transposed im2col of X[nb, conv_L, conv_Cin] (conv wgrad:
 * m = tap*Cin+ci, k = b*L+l).  splits > 1 splits K over grid.z with fp32 atomic accumulation. */
int acb_gemm_ex(const void* A, int a_dtype, const void* B, int b_dtype, float* C, int M, int N, int K, long long sam,
                long long sak, long long sbn, long long sbk, int ldc, int convT, int conv_L, int conv_Cin, int conv_pad,
                int splits, int accumulate, void* stream);
/* tcgen05 weight gradient with MN-major operands (no transposes), fp32 atomic accumulation over row splits:
 *   dW[m, tap*Cin+ci] (+)= sum_{b,l} dY[(b*L+l)*ldy + a_col0 + m] * X[b, l+tap-pad, ci]       (bf16 operands)
 * Linear layers: nb=1, L=rows, taps=1, pad=0.  Convolutions need Cin % 64 == 0. */
int acb_wgrad_bf16(const void* dY, int ldy, int a_col0, int M_out, const void* X, int nb, int L, int Cin, int taps, int pad,
                   long long x_batch_stride, long long x_row_stride, float* dW, int ldc, int accumulate, void* stream);
/* the same launch also producing the bias gradient db[m] (+)= sum_rows dY[row, a_col0 + m]: one extra N tile multiplies the dY
 * tiles with a constant tile of ones, so the column sums cost no second pass over dY (replaces acb_colsum after a Linear /
 * Conv1d backward: nn.Linear / nn.Conv1d bias gradients, e.g. spectranet.py:18-22, timm Mlp fc1/fc2). */
int acb_wgrad_bias_bf16(const void* dY, int ldy, int a_col0, int M_out, const void* X, int nb, int L, int Cin, int taps, int pad,
                        long long x_batch_stride, long long x_row_stride, float* dW, int ldc, int accumulate, float* db, void* stream);
/* weight gradient of a ONE-input-channel Conv1d on its polyphase view (SpectraNet's first block, spectranet.py:18-22 with
 * in_channels = 1): dY is seen as [nb, L, n_groups * ...] rows of n_groups positions, X as overlapping `width`-sample windows of
 * the zero-padded signal;  G[co, n + g*c_group_step] += sum_rows dY[row, a_col0 + g*a_group_stride + co] * X[row, n]  for
 * g < n_groups, co < 64, n < width -- every phase g in one launch, two phases per 128-row tensor-core tile.  G (fp32, caller-zeroed)
 * already points at the column of group 0. */
int acb_wgrad_phases_bf16(const void* dY, int ldy, int a_col0, int n_groups, int a_group_stride, const void* X, int nb, int L,
                          int width, long long x_batch_stride, long long x_row_stride, float* G, int ldc, int c_group_step,
                          void* stream);
/* out[n] (+)= sum_m a[m*ld+n] * (b ? b[m*ld+n] : 1)   (bias / layer-scale gradients; ld <= 0 means N) */
int acb_colsum(const void* a, int a_dtype, const void* b, int b_dtype, long long M, int N, long long ld, float* out,
               int accumulate, void* stream);
int acb_act_fwd(const void* x, int x_dtype, void* y, int y_dtype, int act, long long n, void* stream);
/* dx = dy * act'(x), x = pre-activation */
int acb_act_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, void* dx, int dx_dtype, int act, long long n,
                void* stream);
/* op 0: a+b; 1: a*b; 2: a + g[col]*b; 3: g[col]*a; 4: a*s0 + b*s1; 5: a * g[0] (device scalar) */
int acb_ew(const void* a, int a_dtype, const void* b, int b_dtype, const float* g, void* y, int y_dtype, int op, int C,
           float s0, float s1, long long n, void* stream);
int acb_copy2d(const void* src, int s_dtype, long long lds, void* dst, int d_dtype, long long ldd, long long rows, int cols,
               void* stream);
int acb_gather_cols(const float* X, int ldx, const int* cols, int n, float* Y, long long rows, void* stream);
/* dW[co,ci,tap] = G[(tap*Cin+ci)*Cout + co] (G = convT wgrad GEMM result) */
int acb_unpack_conv_wgrad(const float* G, float* dW, int Cout, int Cin, int k, void* stream);
/* out[ci*(k*Cout) + tap*Cout + co] = w[co,ci,k-1-tap]  (weights of the input-gradient convolution) */
int acb_pack_conv_dgrad_weight(const float* w, void* out, int out_dtype, int Cout, int Cin, int k, void* stream);
/* LayerNorm backward: dx, and dw += sum dy*xhat, db += sum dy (caller zeroes dw/db).  post_act = ACB_ACT_GELU (b = the
 * LayerNorm bias): backward of y = gelu(LN(x)) in one pass, the pre-activation is recomputed from x
 * (SpectraNetBlock norm + GELU, spectranet.py:36-38). */
int acb_layernorm_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, const float* w, const float* b, int post_act,
                      void* dx, int dx_dtype, float* dw, float* db, long long rows, int C, float eps, void* stream);
int acb_attention_varlen_bwd(const void* qkv, int dtype, const void* dout, int dout_dtype, const int* cu_seqlens, int B,
                             int n_heads, int dh, int max_seqlen, float drop_p, long long seed, void* dqkv, int dqkv_dtype,
                             void* stream);
/* grads = [d in_proj.weight (D*7) | d in_proj.bias (D) | dw0 | db0 | dw (D-1) | db (D-1) | d cls_tok (D)] */
int acb_photo_embed_bwd(const float* x, const int* src_idx, int T, int D, const void* dh, int dh_dtype, const float* w,
                        const float* b, float te_drop_p, long long te_seed, float* grads, void* stream);
int acb_scatter_cls(const float* dcls, const int* cu_seqlens, int B, int D, void* dh, int dh_dtype, long long total_tokens,
                    void* stream);
/* depthwise 7x7 without LayerNorm (training forward), flip=1: gradient w.r.t. the input */
int acb_dwconv7(const void* x, int x_dtype, const float* w, const float* bias, int flip, void* y, int y_dtype, int B, int H,
                int W, int C, void* stream);
int acb_dwconv7_wgrad(const void* x, int x_dtype, const void* dy, int dy_dtype, int B, int H, int W, int C, float* dw,
                      float* db, int accumulate, void* stream);
/* 2x2/stride-2 patch gather x[B,H,W,C] -> p[B*(H/2)*(W/2), 4C] (adjoint=0) or its adjoint p -> x (adjoint=1) */
int acb_patch2(const void* x, int x_dtype, void* p, int p_dtype, int B, int H, int W, int C, int adjoint, void* stream);
/* mean over HW (bwd=0: y from x; bwd=1: dx from y) */
int acb_gap(const void* x, int x_dtype, float* y, int B, int HW, int C, int bwd, void* dx, int dx_dtype, void* stream);
/* window 4: MaxPool1d(4) backward; window 0: global max backward (gradient to the first maximum) */
int acb_maxpool_bwd(const void* x, int x_dtype, const void* dy, int dy_dtype, void* dx, int dx_dtype, int B, int L, int C,
                    int window, void* stream);
int acb_moe_combine_bwd(const float* gate, const float* expert_out, const float* dout, float* dgate, float* dexpert_out, int B,
                        int E, int C, void* stream);
/* bwd=0: out = x/||x||; bwd=1: out = d x given dy */
int acb_l2norm(const float* x, const float* dy, float* out, int rows, int C, int bwd, void* stream);
/* focal loss (labels int64, gamma; HyraxBaselineCLS.py:177-191) or soft-target cross entropy (astrominn.py:147,315),
 * mean reduction: loss_out[0] and dlogits = d loss / d logits */
int acb_loss_fwd_bwd(const float* logits, const long long* labels, const float* soft_targets, float gamma, int B, int C,
                     float* loss_out, float* dlogits, void* stream);
int acb_dropout(const void* x, int x_dtype, void* y, int y_dtype, float p, long long seed, long long n, void* stream);
int acb_sumsq(const float* x, long long n, float* out, int accumulate, void* stream);

/* ---- masked-event pre-training (MPTModel, HyraxBaselineCLS.py:194-319) ---------------------------- */
/* _mask_batch (:283-319) on the device: per light curve k = max(int(n_valid*mask_p), 3) tokens, k/3 per band
 * (band = argmax of channels 4:7) drawn uniformly without replacement plus k%3 extras from the remaining valid
 * tokens; channels 2:7 of the chosen rows of x[B,L,7] are zeroed IN PLACE and masked[B,L] set.  The draw is a
 * counter hash of (seed, light curve, token), not torch.randperm: same distribution, different stream. */
int acb_mpt_mask(float* x, const uint8_t* pad, int B, int L, double mask_p, long long seed, uint8_t* masked, void* stream);
/* pred[T,5] = (flux, band logits x3, dt) of the packed tokens (CLS rows ignored); targets are read from the
 * already-masked x exactly as the reference does (:264-272); losses[4] = {lambda_f*L_f*lambda_b*L_b*lambda_dt*L_dt,
 * L_f (MSE), L_b (CE), L_dt (MSE)} as means over masked tokens; dpred[T,5] (optional) = d losses[0] / d pred.
 * workspace: 4 floats. */
int acb_mpt_loss_fwd_bwd(const void* pred, int pred_dtype, const int* src_idx, int T, const float* x, const uint8_t* masked, int L,
                         float lambda_f, float lambda_b, float lambda_dt, float* losses, float* dpred, float* workspace,
                         void* stream);

/* ---- fused optimiser step ------------------------------------------------------------ */
/* One pass over flat fp32 buffers p/g/m/v[n] (16-byte aligned, n % 4 == 0): optional torch-style gradient-norm
 * clip (gnorm_sq = device scalar holding sum(g^2), e.g. from acb_sumsq; coef = min(1, max_norm/(norm+1e-6))),
 * gradient scaling, then Adam (L2 weight decay joins the gradient) or AdamW (decoupled) with per-group
 * hyper-parameters; group i covers [group_end[i-1], group_end[i]) and hyper[6*i..] = {lr, beta1, beta2, eps,
 * weight_decay, decoupled}; both arrays live on the HOST.  step >= 1 is the 1-based step count (bias correction);
 * step_dev (optional, device int) overrides it with a count the device owns, so a captured CUDA graph that increments
 * it replays with the right bias correction.
 * p_bf16 (optional) receives the bf16 copy of the updated weights.  Replaces clip_grad_norm_ + optimizer.step at
 * HyraxBaselineCLS.py:108-120,228,279-280, astrominn.py:151-218,311-326, brew_cider.py:1211. */
int acb_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, int n_groups, const long long* group_end,
                  const float* hyper, int step, const int* step_dev, const float* gnorm_sq, float max_norm, float grad_scale,
                  void* stream);

/* ---- tower groups, training path (ResidualTowerBlock x N in one launch; astrominn.py:44-64,264-300) ---- */
/* ptrs: HOST array, 25 device addresses per tower = {cols (int*, or 0 = columns 0..in-1 of X),
 *   W0, b0, ln1w, ln1b, W1, b1, ln2w, ln2b, W2, b2, Ws, bs,           (parameters, torch (out,in) layout; W0/b0 = 0:
 *                                                                      the pre-GELU start path is supplied in A;
 *                                                                      Ws/bs = 0: identity skip)
 *   gW0, gb0, gln1w, gln1b, gW1, gb1, gln2w, gln2b, gW2, gb2, gWs, gbs} (gradient buffers, backward only: ACCUMULATED)
 * dims: HOST array, 5 ints per tower = {in_dim, hidden, out_dim (<= 32), y_off (column of Y / dY), a_off (column of A / dA)}.
 * fwd: Y[rows, ldy] columns y_off.. = tower(X); A[rows, lda] receives (W0 given) or supplies (W0 = 0) the pre-GELU
 * activations -- the only tensor the backward needs.  Dropout (p = drop_p on both LayerNorm outputs, training only)
 * is a counter hash of (seed, tower, row, column), regenerated in the backward.
 * bwd: parameter gradients accumulated into the g* buffers; dA (optional, layout of A) = d loss / d pre-GELU;
 * dX (optional, [rows, ldx], overwritten) = d loss / d X summed over the towers of the group. */
int acb_tower_group_fwd(const float* X, int ldx, int rows, int n_towers, const long long* ptrs, const int* dims, float* Y, int ldy,
                        float* A, int lda, float drop_p, long long seed, void* stream);
int acb_tower_group_bwd(const float* X, int ldx, int rows, int n_towers, const long long* ptrs, const int* dims, const float* A, int lda,
                        const float* dY, int ldy, float* dA, float* dX, float drop_p, long long seed, void* stream);

/* MultiModalDataset.pad_collate dict format (docs/pre_executed/Fusion_Dataset.ipynb cell 0) -> encoder inputs:
 * events[B,T,Fe] + events_mask[B,T] (nonzero = VALID) -> x[B,T,7] = columns cols[0..6] (dt, dt_prev, logflux,
 * logflux_err, band_ztfg, band_ztfr, band_ztfi of build_event_features, preprocess_multimodal.py:315-336), optional
 * log1p on the two time columns (photo_dataset.py:85-101), channels 0..3 normalised with (x - mean)/(std + 1e-8)
 * (HyraxBaselineCLS.py:157) and pad[B,T] with nonzero = PADDING (the polarity of forward()). */
int acb_collate_events(const float* events, const uint8_t* valid_mask, int B, int T, int Fe, const int* cols, int log1p_dt,
                       const float* mean, const float* stdv, float* x, uint8_t* pad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APPLECIDER_B200_H */
