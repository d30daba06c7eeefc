// Array-level preprocessing kernels: light-curve feature/collate/normalise,
// detection merging + event features, spectrum resampling + mean/MAD scaling, cutout crop + normalise,
// streaming feature statistics.  HBM-bound / latency-bound integer+float work on CUDA cores; index and
// segmentation results are bit-exact, float32 results use the same operation order as the reference.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace {

// ================================ P1: light curves =================================================
// One warp per object: horizon cut (stable compaction), log1p / one-hot features, pad|truncate to max_len,
// mask, (x - mean) / (std + 1e-8) on channels 0..3 of every row including padding.
__global__ void __launch_bounds__(256) prep_lightcurve_kernel(const float* __restrict__ raw, const long long* __restrict__ offsets,
                                                              int B, float horizon, const float* __restrict__ mean,
                                                              const float* __restrict__ stdv, int max_len,
                                                              float* __restrict__ x, uint8_t* __restrict__ mask,
                                                              int* __restrict__ lengths) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const long long r0 = offsets[b];
  const int n = (int)(offsets[b + 1] - r0);
  float mu[4], sd[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    mu[c] = mean[c];
    sd[c] = stdv[c] + 1e-8f;
  }
  float* xb = x + (long long)b * max_len * 7;
  int kept = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    float dt = 0.f, dtp = 0.f, band = 0.f, lf = 0.f, lfe = 0.f;
    bool keep = false;
    if (i < n) {
      const float* r = raw + (r0 + i) * 5;
      dt = r[0]; dtp = r[1]; band = r[2]; lf = r[3]; lfe = r[4];
      keep = dt <= horizon;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int pos = kept + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < max_len) {
      float* o = xb + (long long)pos * 7;
      o[0] = (log1pf(dt) - mu[0]) / sd[0];
      o[1] = (log1pf(dtp) - mu[1]) / sd[1];
      o[2] = (lf - mu[2]) / sd[2];
      o[3] = (lfe - mu[3]) / sd[3];
      const int bi = (int)band;
      o[4] = bi == 0 ? 1.f : 0.f;
      o[5] = bi == 1 ? 1.f : 0.f;
      o[6] = bi == 2 ? 1.f : 0.f;
    }
    kept += __popc(m);
  }
  const int len = min(kept, max_len);
  if (lane == 0 && lengths) lengths[b] = len;
  float pv[7];
#pragma unroll
  for (int c = 0; c < 4; ++c) pv[c] = (0.0f - mu[c]) / sd[c];
  pv[4] = pv[5] = pv[6] = 0.f;
  for (int e = len * 7 + lane; e < max_len * 7; e += 32) xb[e] = pv[e % 7];
  for (int l = lane; l < max_len; l += 32) mask[(long long)b * max_len + l] = l >= len ? 1 : 0;
}

// ================================ P2: merge + event features ======================================
// Per object: per band (reference group order g, i, r) a greedy anchored window merge in fp64 with un-fused multiply/add
// (same rounding sequence as the reference loop), then a stable 3-way merge by time and the float32 feature columns.
// One WARP per object:
//   lanes 0-2  walk "their" band and emit the windows [first, last] (integer / compare work only -- the greedy chain),
//   all lanes  then take windows round-robin and do the fp64 weighted means (the pow() per detection is what costs),
//   all lanes  rank the merged events of the three bands against each other by binary search (stable: ties go to the earlier band),
//   all lanes  write the float32 feature columns.
// Objects with more than P2_MAX_N detections are handled serially by lane 0 (same arithmetic, same results).
constexpr int P2_MAX_N = 512;
constexpr int P2_WARPS = 4;

struct P2Sum { double t, f, e; };

__device__ __forceinline__ P2Sum p2_window(const double* __restrict__ mjd, const double* __restrict__ mag, const double* __restrict__ magerr,
                                           const int* __restrict__ fid, long long i, long long j, int band) {
  const double eps = 1e-8;
  const double c_err = 2.5 / 2.302585092994046;  // 2.5 / ln(10)
  // a 12 h window rarely holds more than a few detections of one band: the first four keep their flux / error (each an fp64 pow,
  // ~150 instructions) in registers for the second pass, later ones are recomputed
  constexpr int KEEP = 4;
  double fl[KEEP], er[KEEP];
  int cnt = 0;
  double totw = 0.0;
  for (long long k = i; k <= j; ++k) {
    if (fid[k] != band) continue;
    const double flux = pow(10.0, -0.4 * (mag[k] - 23.9));
    const double err = __dmul_rn(magerr[k] / c_err, flux);
    totw = __dadd_rn(totw, 1.0 / __dadd_rn(err, eps));
#pragma unroll
    for (int u = 0; u < KEEP; ++u)
      if (cnt == u) { fl[u] = flux; er[u] = err; }
    ++cnt;
  }
  P2Sum s{0.0, 0.0, 0.0};
  cnt = 0;
  for (long long k = i; k <= j; ++k) {
    if (fid[k] != band) continue;
    double flux = 0.0, err = 0.0;
    if (cnt < KEEP) {
#pragma unroll
      for (int u = 0; u < KEEP; ++u)
        if (cnt == u) { flux = fl[u]; err = er[u]; }
    } else {
      flux = pow(10.0, -0.4 * (mag[k] - 23.9));
      err = __dmul_rn(magerr[k] / c_err, flux);
    }
    ++cnt;
    const double w = (1.0 / __dadd_rn(err, eps)) / totw;
    s.t = __dadd_rn(s.t, __dmul_rn(w, mjd[k]));
    s.f = __dadd_rn(s.f, __dmul_rn(w, flux));
    s.e = __dadd_rn(s.e, __dmul_rn(w, err));
  }
  return s;
}

__device__ __forceinline__ void p2_emit(long long dst, double t, double t_first, double t_prev, double f, double e, signed char band,
                                        float* __restrict__ o_dt, float* __restrict__ o_dtp, signed char* __restrict__ o_band,
                                        float* __restrict__ o_lf, float* __restrict__ o_lfe) {
  const float f32 = fmaxf((float)f, 1e-6f);
  o_dt[dst] = (float)(t - t_first);
  o_dtp[dst] = (float)(t - t_prev);
  o_band[dst] = band;
  o_lf[dst] = log10f(f32);
  o_lfe[dst] = (float)(__dmul_rn((double)(float)e, 0.43429448190325176) / (double)f32);
}

// next window of `band` starting the search at detection i: returns false when the band has no detection left
__device__ __forceinline__ bool p2_next_window(const double* __restrict__ mjd, const int* __restrict__ fid, long long r1, int band, double dt_days,
                                               long long& i, long long& j, long long& next) {
  while (i < r1 && fid[i] != band) ++i;
  if (i >= r1) return false;
  const double t0 = mjd[i];
  j = i;
  long long scan = i + 1;
  for (; scan < r1; ++scan) {
    if (fid[scan] != band) continue;
    if (mjd[scan] - t0 <= dt_days) j = scan; else break;
  }
  next = scan;  // first detection of this band beyond the window (or r1)
  return true;
}

__global__ void __launch_bounds__(32 * P2_WARPS) prep_events_kernel(const double* __restrict__ mjd, const double* __restrict__ mag,
                                                          const double* __restrict__ magerr, const int* __restrict__ fid,
                                                          const long long* __restrict__ offsets, int B, double dt_days,
                                                          double* __restrict__ tmp /* [3][total] t,f,e */, signed char* __restrict__ tmp_b,
                                                          long long total, float* __restrict__ o_dt, float* __restrict__ o_dtp,
                                                          signed char* __restrict__ o_band, float* __restrict__ o_lf,
                                                          float* __restrict__ o_lfe, int* __restrict__ n_events) {
  __shared__ unsigned short s_first[P2_WARPS][3][P2_MAX_N];   // window bounds per band (n <= 512), later reused as the merged order
  __shared__ unsigned short s_last[P2_WARPS][3][P2_MAX_N];
  __shared__ int s_cnt[P2_WARPS][4];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int b = blockIdx.x * P2_WARPS + wib;
  if (b >= B) return;
  const long long r0 = offsets[b], r1 = offsets[b + 1];
  const int n = (int)(r1 - r0);
  double* tt = tmp + r0;
  double* tf = tmp + total + r0;
  double* te = tmp + 2 * total + r0;
  signed char* tb = tmp_b + r0;
  const int order[3] = {1, 3, 2};
  if (n > P2_MAX_N) {  // rare: serial path on lane 0
    if (lane != 0) return;
    int seg_start[4];
    int cnt = 0;
    for (int g = 0; g < 3; ++g) {
      seg_start[g] = cnt;
      long long i = r0, j = 0, nx = 0;
      while (p2_next_window(mjd, fid, r1, order[g], dt_days, i, j, nx)) {
        const P2Sum w = p2_window(mjd, mag, magerr, fid, i, j, order[g]);
        tt[cnt] = w.t; tf[cnt] = w.f; te[cnt] = w.e; tb[cnt] = (signed char)(order[g] - 1);
        ++cnt;
        i = nx;
      }
    }
    seg_start[3] = cnt;
    n_events[b] = cnt;
    int p[3] = {seg_start[0], seg_start[1], seg_start[2]};
    double t_first = 0.0, t_prev = 0.0;
    for (int o = 0; o < cnt; ++o) {
      int best = -1;
      for (int g = 0; g < 3; ++g)
        if (p[g] < seg_start[g + 1] && (best < 0 || tt[p[g]] < tt[p[best]])) best = g;
      const int sidx = p[best]++;
      const double t = tt[sidx];
      if (o == 0) { t_first = t; t_prev = t; }
      p2_emit(r0 + o, t, t_first, t_prev, tf[sidx], te[sidx], tb[sidx], o_dt, o_dtp, o_band, o_lf, o_lfe);
      t_prev = t;
    }
    return;
  }
  // ---- A: the three greedy window chains, one lane per band ----
  if (lane < 3) {
    int c = 0;
    long long i = r0, j = 0, nx = 0;
    while (p2_next_window(mjd, fid, r1, order[lane], dt_days, i, j, nx)) {
      s_first[wib][lane][c] = (unsigned short)(i - r0);
      s_last[wib][lane][c] = (unsigned short)(j - r0);
      ++c;
      i = nx;
    }
    s_cnt[wib][lane] = c;
  }
  __syncwarp();
  const int c0 = s_cnt[wib][0], c1 = s_cnt[wib][1], c2 = s_cnt[wib][2];
  const int seg[4] = {0, c0, c0 + c1, c0 + c1 + c2};
  const int cnt = seg[3];
  if (lane == 0) n_events[b] = cnt;
  // ---- B: weighted means of the windows (fp64, the reference's summation order inside a window) ----
  for (int sidx = lane; sidx < cnt; sidx += 32) {
    const int g = sidx < seg[1] ? 0 : (sidx < seg[2] ? 1 : 2);
    const int k = sidx - seg[g];
    const P2Sum w = p2_window(mjd, mag, magerr, fid, r0 + s_first[wib][g][k], r0 + s_last[wib][g][k], order[g]);
    tt[sidx] = w.t; tf[sidx] = w.f; te[sidx] = w.e; tb[sidx] = (signed char)(order[g] - 1);
  }
  __syncwarp();
  // ---- C: stable 3-way merge by rank: position = index in own band + events of the other bands that come first ----
  unsigned short* ord = &s_first[wib][0][0];  // 3 * P2_MAX_N entries >= cnt
  __syncwarp();
  for (int sidx = lane; sidx < cnt; sidx += 32) {
    const int g = sidx < seg[1] ? 0 : (sidx < seg[2] ? 1 : 2);
    const double t = tt[sidx];
    int pos = sidx - seg[g];
    for (int g2 = 0; g2 < 3; ++g2) {
      if (g2 == g) continue;
      // events of band g2 placed before this one: time < t, or time == t when g2 is the earlier band (the serial merge is stable)
      int lo = seg[g2], hi = seg[g2 + 1];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double tm = tt[mid];
        if (tm < t || (tm == t && g2 < g)) lo = mid + 1; else hi = mid;
      }
      pos += lo - seg[g2];
    }
    // ord is written after every lane has finished reading s_first / s_last in phase B (the __syncwarp above)
    ord[pos] = (unsigned short)sidx;
  }
  __syncwarp();
  // ---- D: feature columns ----
  const double t_first = cnt > 0 ? tt[ord[0]] : 0.0;
  for (int o = lane; o < cnt; o += 32) {
    const int sidx = ord[o];
    const double t = tt[sidx];
    const double t_prev = o > 0 ? tt[ord[o - 1]] : t;
    p2_emit(r0 + o, t, t_first, t_prev, tf[sidx], te[sidx], tb[sidx], o_dt, o_dtp, o_band, o_lf, o_lfe);
  }
}

// ================================ selection helper ===================================================
// k-th smallest (0-based) of `n` keys produced by key(i), radix select over NB-bit unsigned keys, 8 bits/pass.
// k-th smallest key among key(0..n-1) by MSB-first radix selection with 8-bit digits, for blocks of exactly 256 threads:
//   * the 256-bin histogram is scanned by all threads (one bin each: warp scan + cross-warp offsets), not by thread 0;
//   * as soon as the bin that holds rank k has <= 64 members the pass loop stops: the members are gathered into shared memory
//     and ranked against each other (one candidate per thread) -- smooth spectra / sky-dominated cutouts share their sign,
//     exponent and leading mantissa bits, so this happens after 2-3 of the 4 (fp32) or 8 (fp64) passes.
// hist: 256 counters; sh_k: 8 words of scratch; cand: 257 keys (slot 256 receives the result).
constexpr int SEL_THREADS = 256;
constexpr int SEL_CAND = 64;  // candidates ranked against each other: 64 x 64 compares cost less than one more histogram pass
template <typename KeyT, typename KeyFn>
__device__ KeyT block_select(int n, int k, KeyFn key, unsigned* hist /* 256 */, unsigned* sh_k /* 8 */, KeyT* cand /* 257 */) {
  constexpr int NBITS = sizeof(KeyT) * 8;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  KeyT prefix = 0, pmask = 0;
  for (int shift = NBITS - 8; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) {  // (warp-aggregating these atomics with match.any was measured 25 % SLOWER)
      const KeyT kk = key(i);
      if ((kk & pmask) == prefix) atomicAdd(&hist[(unsigned)((kk >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    // inclusive scan of the 256 bins, one per thread
    const unsigned v = hist[tid];
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sh_k[wid] = inc;  // 8 warp totals
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < wid; ++w) base += sh_k[w];
    const unsigned excl = base + inc - v;
    __syncthreads();
    if (excl <= (unsigned)k && (unsigned)k < excl + v) {  // exactly one thread: the bin holding rank k
      sh_k[0] = (unsigned)tid;
      sh_k[1] = (unsigned)k - excl;
      sh_k[2] = v;
    }
    __syncthreads();
    prefix |= (KeyT)sh_k[0] << shift;
    pmask |= (KeyT)0xff << shift;
    k = (int)sh_k[1];
    const unsigned members = sh_k[2];
    __syncthreads();
    if (shift > 0 && members <= (unsigned)SEL_CAND) {
      if (tid == 0) sh_k[3] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += SEL_THREADS) {
        const KeyT kk = key(i);
        if ((kk & pmask) == prefix) cand[atomicAdd(&sh_k[3], 1u)] = kk;
      }
      __syncthreads();
      if (tid < (int)members) {
        const KeyT c = cand[tid];
        int rank = 0;
        for (int j = 0; j < (int)members; ++j) {
          const KeyT o = cand[j];
          rank += (o < c) || (o == c && j < tid);
        }
        if (rank == k) cand[SEL_THREADS] = c;  // slot 256: the result (exactly one thread has this rank)
      }
      __syncthreads();
      const KeyT res = cand[SEL_THREADS];
      __syncthreads();
      return res;
    }
  }
  return prefix;
}

__device__ __forceinline__ unsigned long long dkey(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ unsigned fkey(float v) {
  const unsigned u = __float_as_uint(v);
  return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
  const unsigned u = (k >> 31) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

__device__ double block_sum_d(double v, double* sh /* 33 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  return sh[32];
}

// ================================ P3: spectrum resampling ============================================
// One CTA per spectrum: finite filter (stable), sort by wavelength when needed (bitonic, smem), linear
// interpolation with linear extrapolation at searchsorted-left intervals (un-fused fp64, scipy order of
// operations), mean, MAD by radix selection, (y - mean) / scale -> float32.
__global__ void __launch_bounds__(256) prep_spectrum_kernel(const double* __restrict__ wl, const double* __restrict__ fx,
                                                            const long long* __restrict__ offsets, int cap,
                                                            const float* __restrict__ grid, int n_grid, float* __restrict__ out,
                                                            int* __restrict__ idx_out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* xs = reinterpret_cast<double*>(sm_raw);
  double* ys = xs + cap;
  double* yg = ys + cap;
  double* red = yg + n_grid;                                  // 33 doubles
  unsigned* hist = reinterpret_cast<unsigned*>(red + 34);     // 256
  unsigned* shk = hist + 256;                                  // 8
  int* shi = reinterpret_cast<int*>(shk + 8);                  // 16 counters
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(shi + 16);  // 257 keys (8-byte aligned: 34*8 + 280*4 bytes before)
  const int b = blockIdx.x;
  const long long r0 = offsets[b];
  const int n_in = (int)(offsets[b + 1] - r0);
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  float* ob = out + (long long)b * n_grid;

  // ---- stable compaction of finite samples ----
  if (tid == 0) shi[0] = 0;
  __syncthreads();
  for (int base = 0; base < n_in; base += nthr) {
    const int i = base + tid;
    double xv = 0.0, yv = 0.0;
    bool ok = false;
    if (i < n_in) {
      xv = wl[r0 + i];
      yv = fx[r0 + i];
      ok = isfinite(xv) && isfinite(yv);
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    int* wcnt = shi + 1;  // per-warp counts
    if (lane == 0) wcnt[wid] = __popc(m);
    __syncthreads();
    int off = shi[0];
    for (int w = 0; w < wid; ++w) off += wcnt[w];
    if (ok) {
      const int pos = off + __popc(m & ((1u << lane) - 1u));
      xs[pos] = xv;
      ys[pos] = yv;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < nw; ++w) t += wcnt[w];
      shi[0] += t;
    }
    __syncthreads();
  }
  const int n = shi[0];
  if (n < 2) {
    for (int g = tid; g < n_grid; g += nthr) {
      ob[g] = CUDART_NAN_F;
      if (idx_out) idx_out[(long long)b * n_grid + g] = -1;
    }
    return;
  }
  // ---- sort by wavelength if needed ----
  int unsorted = 0;
  for (int i = tid; i + 1 < n; i += nthr) unsorted |= (xs[i] > xs[i + 1]);
  unsorted = __syncthreads_or(unsorted);
  if (unsorted) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = n + tid; i < np2; i += nthr) { xs[i] = CUDART_INF; ys[i] = 0.0; }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += nthr) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool up = (i & k) == 0;
            const double a = xs[i], c = xs[ixj];
            if ((a > c) == up) {
              xs[i] = c; xs[ixj] = a;
              const double t = ys[i]; ys[i] = ys[ixj]; ys[ixj] = t;
            }
          }
        }
        __syncthreads();
      }
    }
  }
  // ---- interpolation ----
  // every thread owns a CONTIGUOUS run of grid points: the wavelength grid ascends, so the searchsorted position of a point
  // is found by an exponential + binary search that starts at the previous point's position (a few steps instead of log2(n))
  double s_loc = 0.0;
  const int per = (n_grid + nthr - 1) / nthr;
  int prev_lo = 0;
  double prev_x = -CUDART_INF;
  for (int g = tid * per; g < min(n_grid, (tid + 1) * per); ++g) {
    const double xn = (double)grid[g];
    int lo = xn >= prev_x ? prev_lo : 0;  // first index with xs[idx] >= xn  (numpy searchsorted side='left')
    int hi = lo, step = 1;
    while (hi < n && xs[hi] < xn) { lo = hi + 1; hi += step; step <<= 1; }
    hi = min(hi, n);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (xs[mid] < xn) lo = mid + 1; else hi = mid;
    }
    prev_lo = lo;
    prev_x = xn;
    if (idx_out) idx_out[(long long)b * n_grid + g] = lo;
    int ih = min(max(lo, 1), n - 1);
    const int il = ih - 1;
    const double slope = __ddiv_rn(__dsub_rn(ys[ih], ys[il]), __dsub_rn(xs[ih], xs[il]));
    const double v = __dadd_rn(__dmul_rn(slope, __dsub_rn(xn, xs[il])), ys[il]);
    yg[g] = v;
    if (!isnan(v)) s_loc += v;
  }
  __syncthreads();
  int nfin_loc = 0;
  for (int g = tid; g < n_grid; g += nthr) nfin_loc += !isnan(yg[g]);
  const int nfin = (int)(block_sum_d((double)nfin_loc, red) + 0.5);
  const double mean = block_sum_d(s_loc, red) / (double)nfin;
  // ---- median and MAD (NaNs sort last: select among the nfin smallest keys) ----
  double scale = 1.0;
  if (nfin > 0) {
    auto key_y = [&](int i) { const double v = yg[i]; return isnan(v) ? ~0ull : dkey(v); };
    double med = dkey_inv(block_select<unsigned long long>(n_grid, (nfin - 1) / 2, key_y, hist, shk, cand));
    if ((nfin & 1) == 0) med = 0.5 * (med + dkey_inv(block_select<unsigned long long>(n_grid, nfin / 2, key_y, hist, shk, cand)));
    auto key_d = [&](int i) { const double v = yg[i]; return isnan(v) ? ~0ull : dkey(fabs(v - med)); };
    double mad = dkey_inv(block_select<unsigned long long>(n_grid, (nfin - 1) / 2, key_d, hist, shk, cand));
    if ((nfin & 1) == 0) mad = 0.5 * (mad + dkey_inv(block_select<unsigned long long>(n_grid, nfin / 2, key_d, hist, shk, cand)));
    if (!isfinite(mad) || mad == 0.0) {
      double q = 0.0;
      for (int g = tid; g < n_grid; g += nthr) {
        const double v = yg[g];
        if (!isnan(v)) q += (v - mean) * (v - mean);
      }
      const double sd = sqrt(block_sum_d(q, red) / (double)nfin);
      scale = (isfinite(sd) && sd > 0.0) ? sd : 1.0;
    } else {
      scale = mad;
    }
  }
  for (int g = tid; g < n_grid; g += nthr) ob[g] = (float)((yg[g] - mean) / scale);
}

// ---- register-resident variant (n_grid <= 14 x 256, i.e. the reference's 3481-point grid) ----------------------------------------
// Same arithmetic as prep_spectrum_kernel, different data placement: a thread keeps the 14 resampled values of its contiguous run
// of grid points in REGISTERS (the radix selects then read no shared memory and the 28 KB `yg` array disappears: 3 CTAs per SM
// instead of 2), an all-finite spectrum is copied in with independent loads instead of the barrier-per-chunk compaction loop, a
// radix pass costs two barriers (warp 0 scans the bins while the others zero the next pass's histogram), and the final
// (y - mean) / scale is a Markstein-corrected multiplication by the reciprocal (correctly rounded like the division it replaces).
constexpr int SP_PER = 14;

__device__ __forceinline__ unsigned long long spectrum_select_reg(const unsigned long long (&key)[SP_PER], int k, unsigned* hist /* 2 x 256 */,
                                                                  unsigned* sh_k /* 8 */, unsigned long long* cand /* 257 */) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  hist[tid] = 0;
  if (tid == 0) sh_k[3] = 0;
  __syncthreads();
  unsigned long long prefix = 0, pmask = 0;
  int set = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    unsigned* h = hist + set * 256;
    if (shift == 56) {  // sign + leading exponent bits: usually ONE bin for the whole spectrum
      const unsigned d0 = __shfl_sync(0xffffffffu, (unsigned)(key[0] >> 56), 0);
      bool same = true;
#pragma unroll
      for (int e = 0; e < SP_PER; ++e) same &= (unsigned)(key[e] >> 56) == d0;
      if (__all_sync(0xffffffffu, same)) {
        if (lane == 0) atomicAdd(&h[d0], 32u * SP_PER);
      } else {
#pragma unroll
        for (int e = 0; e < SP_PER; ++e) atomicAdd(&h[(unsigned)(key[e] >> 56)], 1u);
      }
    } else {
#pragma unroll
      for (int e = 0; e < SP_PER; ++e)
        if ((key[e] & pmask) == prefix) atomicAdd(&h[(unsigned)(key[e] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(set ^ 1) * 256 + tid] = 0;  // the next pass's histogram
    if (wid == 0) {
      const uint4 lo = reinterpret_cast<const uint4*>(h)[2 * lane], hi = reinterpret_cast<const uint4*>(h)[2 * lane + 1];
      const unsigned c[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      const unsigned tot = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7];
      unsigned inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      unsigned kk = (unsigned)k - (inc - tot);
      if (kk < tot) {  // exactly one lane
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (kk < c[j]) {
            sh_k[0] = (unsigned)(8 * lane + j);
            sh_k[1] = kk;
            sh_k[2] = c[j];
            kk = 0xffffffffu;
          } else {
            kk -= c[j];
          }
        }
      }
    }
    __syncthreads();
    prefix |= (unsigned long long)sh_k[0] << shift;
    pmask |= 0xffull << shift;
    k = (int)sh_k[1];
    const unsigned members = sh_k[2];
    set ^= 1;
    if (shift > 0 && members <= (unsigned)SEL_CAND) {
#pragma unroll
      for (int e = 0; e < SP_PER; ++e)
        if ((key[e] & pmask) == prefix) cand[atomicAdd(&sh_k[3], 1u)] = key[e];
      __syncthreads();
      if (tid < (int)members) {
        const unsigned long long c = cand[tid];
        int rank = 0;
        for (int j = 0; j < (int)members; ++j) {
          const unsigned long long o = cand[j];
          rank += (o < c) || (o == c && j < tid);
        }
        if (rank == k) cand[SEL_THREADS] = c;
      }
      __syncthreads();
      return cand[SEL_THREADS];
    }
  }
  return prefix;
}

__global__ void __launch_bounds__(256, 4) prep_spectrum_reg_kernel(const double* __restrict__ wl, const double* __restrict__ fx,
                                                                   const long long* __restrict__ offsets, int cap,
                                                                   const float* __restrict__ grid, int n_grid, float* __restrict__ out,
                                                                   int* __restrict__ idx_out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* xs = reinterpret_cast<double*>(sm_raw);
  double* ys = xs + cap;
  __shared__ double red[34];
  __shared__ __align__(16) unsigned hist[512];
  __shared__ unsigned shk[8];
  __shared__ int shi[16];
  __shared__ unsigned long long cand[SEL_THREADS + 1];
  const int b = blockIdx.x;
  const long long r0 = offsets[b];
  const int n_in = (int)(offsets[b + 1] - r0);
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  float* ob = out + (long long)b * n_grid;

  // ---- load: straight copy when every sample is finite (the usual case), else the stable compaction of the finite samples ----
  bool bad = false;
  for (int i = tid; i < n_in; i += nthr) {
    const double xv = __ldcs(wl + r0 + i), yv = __ldcs(fx + r0 + i);
    xs[i] = xv;
    ys[i] = yv;
    bad |= !(isfinite(xv) && isfinite(yv));
  }
  int n = n_in;
  if (__syncthreads_or(bad)) {
    if (tid == 0) shi[0] = 0;
    __syncthreads();
    for (int base = 0; base < n_in; base += nthr) {
      const int i = base + tid;
      double xv = 0.0, yv = 0.0;
      bool ok = false;
      if (i < n_in) {  // in place: position <= i, and this chunk is in registers before anything of it is overwritten
        xv = xs[i];
        yv = ys[i];
        ok = isfinite(xv) && isfinite(yv);
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      int* wcnt = shi + 1;
      if (lane == 0) wcnt[wid] = __popc(m);
      __syncthreads();
      int off = shi[0];
      for (int w = 0; w < wid; ++w) off += wcnt[w];
      if (ok) {
        const int pos = off + __popc(m & ((1u << lane) - 1u));
        xs[pos] = xv;
        ys[pos] = yv;
      }
      __syncthreads();
      if (tid == 0) {
        int t = 0;
        for (int w = 0; w < nw; ++w) t += wcnt[w];
        shi[0] += t;
      }
      __syncthreads();
    }
    n = shi[0];
  }
  if (n < 2) {
    for (int g = tid; g < n_grid; g += nthr) {
      ob[g] = CUDART_NAN_F;
      if (idx_out) idx_out[(long long)b * n_grid + g] = -1;
    }
    return;
  }
  // ---- sort by wavelength if needed ----
  int unsorted = 0;
  for (int i = tid; i + 1 < n; i += nthr) unsorted |= (xs[i] > xs[i + 1]);
  unsorted = __syncthreads_or(unsorted);
  if (unsorted) {
    // bitonic network in its all-ascending form (first step of every merge pairs i with i ^ (2k - 1), the rest with i ^ j): virtual
    // +inf padding above n never has to move, so pairs with a partner >= n are skipped and n need not be a power of two
    for (int k = 2; (k >> 1) < n; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n; i += nthr) {
          const int ixj = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
          if (ixj > i && ixj < n) {
            const double a = xs[i], c = xs[ixj];
            if (a > c) {
              xs[i] = c; xs[ixj] = a;
              const double t = ys[i]; ys[i] = ys[ixj]; ys[ixj] = t;
            }
          }
        }
        __syncthreads();
      }
    }
  }
// ---- interpolation: thread t owns grid points [14 t, 14 t + 14); galloping + binary search from the previous position ----
  double yv[SP_PER];
  double s_loc = 0.0;
  int nfin_loc = 0;
  {
    int prev_lo = 0, prev_ih = -1;
    double prev_x = -CUDART_INF, slope = 0.0, x_il = 0.0, y_il = 0.0;
#pragma unroll
    for (int e = 0; e < SP_PER; ++e) {
      const int g = tid * SP_PER + e;
      yv[e] = CUDART_NAN;
      if (g < n_grid) {
        const double xn = (double)grid[g];
        int lo = xn >= prev_x ? prev_lo : 0;  // first index with xs[idx] >= xn  (numpy searchsorted side='left')
        int hi = lo, step = 1;
        while (hi < n && xs[hi] < xn) { lo = hi + 1; hi += step; step <<= 1; }
        hi = min(hi, n);
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (xs[mid] < xn) lo = mid + 1; else hi = mid;
        }
        prev_lo = lo;
        prev_x = xn;
        if (idx_out) idx_out[(long long)b * n_grid + g] = lo;
        const int ih = min(max(lo, 1), n - 1);
        const int il = ih - 1;
        if (ih != prev_ih) {  // consecutive grid points usually share the interval: one fp64 division per interval, not per point
          x_il = xs[il];
          y_il = ys[il];
          slope = __ddiv_rn(__dsub_rn(ys[ih], y_il), __dsub_rn(xs[ih], x_il));
          prev_ih = ih;
        }
        const double v = __dadd_rn(__dmul_rn(slope, __dsub_rn(xn, x_il)), y_il);
        yv[e] = v;
        if (!isnan(v)) { s_loc += v; ++nfin_loc; }
      }
    }
  }
  const int nfin = (int)(block_sum_d((double)nfin_loc, red) + 0.5);
  const double mean = block_sum_d(s_loc, red) / (double)nfin;
  // ---- median and MAD (NaNs and the slots past the grid carry the largest key: select among the nfin smallest) ----
  double scale = 1.0;
  if (nfin > 0) {
    unsigned long long key[SP_PER];
#pragma unroll
    for (int e = 0; e < SP_PER; ++e) key[e] = isnan(yv[e]) ? ~0ull : dkey(yv[e]);
    double med = dkey_inv(spectrum_select_reg(key, (nfin - 1) / 2, hist, shk, cand));
    if ((nfin & 1) == 0) med = 0.5 * (med + dkey_inv(spectrum_select_reg(key, nfin / 2, hist, shk, cand)));
#pragma unroll
    for (int e = 0; e < SP_PER; ++e) key[e] = isnan(yv[e]) ? ~0ull : dkey(fabs(yv[e] - med));
    double mad = dkey_inv(spectrum_select_reg(key, (nfin - 1) / 2, hist, shk, cand));
    if ((nfin & 1) == 0) mad = 0.5 * (mad + dkey_inv(spectrum_select_reg(key, nfin / 2, hist, shk, cand)));
    if (!isfinite(mad) || mad == 0.0) {
      double q = 0.0;
#pragma unroll
      for (int e = 0; e < SP_PER; ++e)
        if (!isnan(yv[e])) q += (yv[e] - mean) * (yv[e] - mean);
      const double sd = sqrt(block_sum_d(q, red) / (double)nfin);
      scale = (isfinite(sd) && sd > 0.0) ? sd : 1.0;
    } else {
      scale = mad;
    }
  }
  // ---- (y - mean) / scale -> float32, staged through shared memory (xs is free now) so that the rows leave as full lines ----
  __syncthreads();
  float* stage = reinterpret_cast<float*>(xs);
  const double rinv = 1.0 / scale;
#pragma unroll
  for (int e = 0; e < SP_PER; ++e) {
    const int g = tid * SP_PER + e;
    if (g < n_grid) {
      const double num = yv[e] - mean;
      const double q0 = num * rinv;
      double q = fma(fma(-q0, scale, num), rinv, q0);  // correctly rounded num / scale (Markstein) ...
      if (!isfinite(q0) || !isfinite(rinv)) q = num / scale;  // ... except at the edges of the range
      stage[g] = (float)q;
    }
  }
  __syncthreads();
  for (int g = tid; g < n_grid; g += nthr) ob[g] = stage[g];
}

// ================================ P4: cutouts =========================================================
// mode 0: per channel  x -= lower_median; x /= (unbiased std + 1e-8)      (ImageAndMetadataDataset.get_image)
// mode 2: per channel  x -= median; x /= population std (<= 1e-8 -> 1)     (Fusion_Dataset._normalize_image)
// CTA per (image, channel); the plane lives in shared memory, the median comes from a radix select.
__global__ void __launch_bounds__(256) cutout_median_kernel(const float* __restrict__ img, int C, int H, int W, int i1, int S,
                                                            int mode, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* pl = reinterpret_cast<float*>(sm_raw);
  double* red = reinterpret_cast<double*>(pl + ((S * S + 3) & ~3));
  unsigned* hist = reinterpret_cast<unsigned*>(red + 34);
  unsigned* shk = hist + 256;   // 8
  unsigned* cand = shk + 8;     // 257
  const int bc = blockIdx.x;  // b*C + c
  const float* src = img + (long long)bc * H * W;
  const int n = S * S;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int yy = i / S, xx = i - yy * S;
    pl[i] = src[(yy + i1) * W + (xx + i1)];
  }
  __syncthreads();
  auto key = [&](int i) { return fkey(pl[i]); };
  float med = fkey_inv(block_select<unsigned>(n, (n - 1) / 2, key, hist, shk, cand));
  if (mode == 2 && (n & 1) == 0) med = 0.5f * (med + fkey_inv(block_select<unsigned>(n, n / 2, key, hist, shk, cand)));
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float p = pl[i] - med;
    pl[i] = p;
    s += (double)p;
  }
  const double mean = block_sum_d(s, red) / (double)n;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)pl[i] - mean;
    q += d * d;
  }
  q = block_sum_d(q, red);
  float denom;
  if (mode == 0) {
    denom = (float)sqrt(q / (double)(n - 1)) + 1e-8f;
  } else {
    const double sd = sqrt(q / (double)n);
    denom = (isfinite(sd) && sd > 1e-8) ? (float)sd : 1.0f;
  }
  float* dst = out + (long long)bc * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = pl[i] / denom;
}

// Register-resident variant for planes of <= 16 x 256 pixels (the 63 x 63 ZTF cutout and every crop of it): each thread keeps its
// 16 order-preserving keys in registers, so a radix pass is 4 instructions per pixel with no shared-memory reads, the statistics
// are ONE fp64 pass (sum and sum of squares of the median-centred pixels) and the only shared memory is the 256-bin histogram.
// Pixels past n carry the key 0xffffffff, which sorts after every real value and therefore never changes a rank below n.
constexpr int CUT_EPT = 16;

// hist: 4 x 256 counters (one set per radix pass) and sh_k[3], all ZERO on entry (the caller zeroes them before a barrier): a pass
// then costs two barriers -- histogram complete / owner bin published -- and warp 0 alone scans the 256 bins (8 per lane).
__device__ __forceinline__ unsigned cutout_select_reg(const unsigned (&key)[CUT_EPT], int k, unsigned* hist, unsigned* sh_k, unsigned* cand) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned prefix = 0, pmask = 0;
  for (int shift = 24; shift >= 0; shift -= 8, hist += 256) {
    if (shift == 24) {
      // sign + leading exponent bits: a sky-dominated plane puts nearly every pixel into ONE bin -- a warp whose 512 digits are
      // all equal adds them with one atomic
      const unsigned d0 = __shfl_sync(0xffffffffu, key[0] >> 24, 0);
      bool same = true;
#pragma unroll
      for (int e = 0; e < CUT_EPT; ++e) same &= (key[e] >> 24) == d0;
      if (__all_sync(0xffffffffu, same)) {
        if (lane == 0) atomicAdd(&hist[d0], 32u * CUT_EPT);
      } else {
#pragma unroll
        for (int e = 0; e < CUT_EPT; ++e) atomicAdd(&hist[key[e] >> 24], 1u);
      }
    } else {
#pragma unroll
      for (int e = 0; e < CUT_EPT; ++e)
        if ((key[e] & pmask) == prefix) atomicAdd(&hist[(key[e] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (wid == 0) {
      const uint4 lo = reinterpret_cast<const uint4*>(hist)[2 * lane], hi = reinterpret_cast<const uint4*>(hist)[2 * lane + 1];
      const unsigned c[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      const unsigned tot = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7];
      unsigned inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      unsigned kk = (unsigned)k - (inc - tot);  // rank inside this lane's 8 bins (wraps when the rank is in an earlier lane)
      if (kk < tot) {                           // exactly one lane
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (kk < c[j]) {
            sh_k[0] = (unsigned)(8 * lane + j);
            sh_k[1] = kk;
            sh_k[2] = c[j];
            kk = 0xffffffffu;
          } else {
            kk -= c[j];
          }
        }
      }
    }
    __syncthreads();
    prefix |= sh_k[0] << shift;
    pmask |= 0xffu << shift;
    k = (int)sh_k[1];
    const unsigned members = sh_k[2];  // (the next pass rewrites sh_k only after its histogram barrier)
    if (shift > 0 && members <= (unsigned)SEL_CAND) {
#pragma unroll
      for (int e = 0; e < CUT_EPT; ++e)
        if ((key[e] & pmask) == prefix) cand[atomicAdd(&sh_k[3], 1u)] = key[e];
      __syncthreads();
      if (tid < (int)members) {
        const unsigned c = cand[tid];
        int rank = 0;
        for (int j = 0; j < (int)members; ++j) {
          const unsigned o = cand[j];
          rank += (o < c) || (o == c && j < tid);
        }
        if (rank == k) cand[SEL_THREADS] = c;
      }
      __syncthreads();
      return cand[SEL_THREADS];
    }
  }
  return prefix;
}

// Persistent CTAs (5 per SM): while a plane is being processed its successor is already in flight -- each thread copies ITS 16 pixels
// of the next plane into a 16 KB staging buffer with 4-byte cp.async (the plane start is only 4-byte aligned) as soon as it has
// moved the current ones into registers, so no barrier is involved and the DRAM latency hides behind a whole plane of work.
__device__ __forceinline__ void cutout_cp_async4(const float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(256, 5) cutout_median_reg_kernel(const float* __restrict__ img, int planes, int H, int W, int i1, int S,
                                                                int mode, float* __restrict__ out) {
  __shared__ __align__(16) unsigned hist[4 * 256];
  __shared__ unsigned shk[8];
  __shared__ unsigned cand[SEL_THREADS + 1];
  __shared__ double red[2][9];
  __shared__ float stage[CUT_EPT * 256];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = S * S;
  const double inv_n = 1.0 / (double)n, inv_nm1 = 1.0 / (double)(n - 1);
  const bool crop = S != W;
  auto prefetch = [&](int plane) {
    const float* src = img + (long long)plane * H * W;
    if (!crop) {
#pragma unroll
      for (int e = 0; e < CUT_EPT; ++e)
        if (tid + e * 256 < n) cutout_cp_async4(&stage[tid + e * 256], src + tid + e * 256);
    } else {
      // pixel tid + 256 e of the cropped plane: (row, column) advance by (256 / S, 256 % S) per step -- no division per pixel
      const int qs = 256 / S, rs = 256 - qs * S;
      int yy = tid / S, xx = tid - yy * S;
#pragma unroll
      for (int e = 0; e < CUT_EPT; ++e) {
        if (tid + e * 256 < n) cutout_cp_async4(&stage[tid + e * 256], src + (yy + i1) * W + (xx + i1));
        xx += rs;
        yy += qs;
        if (xx >= S) {
          xx -= S;
          ++yy;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int plane = blockIdx.x;
  if (plane < planes) prefetch(plane);
  for (; plane < planes; plane += gridDim.x) {
    // (the barriers of the previous plane's statistics stand between its last readers of hist / shk / cand / red and these writes)
#pragma unroll
    for (int j = 0; j < 4; ++j) hist[tid + 256 * j] = 0;
    if (tid < 8) shk[tid] = 0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    unsigned key[CUT_EPT];
#pragma unroll
    for (int e = 0; e < CUT_EPT; ++e) key[e] = tid + e * 256 < n ? fkey(stage[tid + e * 256]) : 0xffffffffu;
    if (plane + (int)gridDim.x < planes) prefetch(plane + (int)gridDim.x);  // a thread refills only the slots it has just read
    __syncthreads();
    float med = fkey_inv(cutout_select_reg(key, (n - 1) / 2, hist, shk, cand));
    if (mode == 2 && (n & 1) == 0) {  // np.median of an even count: mean of the two middle values
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) hist[tid + 256 * j] = 0;
      if (tid < 8) shk[tid] = 0;
      __syncthreads();
      med = 0.5f * (med + fkey_inv(cutout_select_reg(key, n / 2, hist, shk, cand)));
    }
    // n > 15 x 256 (the 63 x 63 plane and its neighbours): only a thread's LAST pixel can lie past the plane
    const bool full = n > (CUT_EPT - 1) * 256;
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int e = 0; e < CUT_EPT; ++e) {
      key[e] = __float_as_uint(fkey_inv(key[e]) - med);  // the registers now hold the median-centred pixels
      if ((full && e < CUT_EPT - 1) || tid + e * 256 < n) {
        const double d = (double)__uint_as_float(key[e]);
        s += d;
        q = fma(d, d, q);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      red[0][wid] = s;
      red[1][wid] = q;
    }
    __syncthreads();
    s = 0.0, q = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {  // every thread adds the 8 warp partials in the same order
      s += red[0][w];
      q += red[1][w];
    }
    // sum of squares about the mean; the pixels are already centred on the median, so the subtraction loses nothing in fp64
    const double ss = fmax(q - s * s * inv_n, 0.0);
    float denom;
    if (mode == 0) {
      denom = (float)sqrt(ss * inv_nm1) + 1e-8f;
    } else {
      const double sd = sqrt(ss * inv_n);
      denom = (isfinite(sd) && sd > 1e-8) ? (float)sd : 1.0f;
    }
    // p / denom, correctly rounded without the 20-instruction division sequence: with r = RN(1 / denom), q0 = RN(p r) and the
    // exact remainder p - q0 denom (one fma), RN(q0 + rem r) is the correctly rounded quotient (Markstein); denom >= 1e-8 is normal
    const float r = __frcp_rn(denom);
    float* dst = out + (long long)plane * n;
#pragma unroll
    for (int e = 0; e < CUT_EPT; ++e) {
      const int i = tid + e * 256;
      if ((full && e < CUT_EPT - 1) || i < n) {
        const float pe = __uint_as_float(key[e]);
        const float q0 = pe * r;
        dst[i] = fmaf(fmaf(-q0, denom, pe), r, q0);
      }
    }
  }
}

// mode 1: x / ||x||_2 over all channels of the (cropped) cutout.  CTA per image.
__global__ void __launch_bounds__(256) cutout_l2_kernel(const float* __restrict__ img, int C, int H, int W, int i1, int S,
                                                        float* __restrict__ out) {
  __shared__ double red[34];
  const int b = blockIdx.x;
  const int n = C * S * S;
  const float* src = img + (long long)b * C * H * W;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / (S * S), rem = i - c * S * S, yy = rem / S, xx = rem - yy * S;
    const double v = (double)src[(c * H + yy + i1) * W + xx + i1];
    s += v * v;
  }
  const float nrm = (float)sqrt(block_sum_d(s, red));
  float* dst = out + (long long)b * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / (S * S), rem = i - c * S * S, yy = rem / S, xx = rem - yy * S;
    dst[i] = src[(c * H + yy + i1) * W + xx + i1] / nrm;
  }
}

// ================================ P5: feature statistics =============================================
__global__ void __launch_bounds__(256) feature_sums_kernel(const float* __restrict__ data, long long rows, int F, int rows_per_block,
                                                           double* __restrict__ sums /* [2F] */) {
  __shared__ double sh[2 * 256];
  const int active = (256 / F) * F;
  for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  if (threadIdx.x < active && r0 < r1) {
    const int col = threadIdx.x % F;
    double s = 0.0, q = 0.0;
    const long long e1 = (r1 - r0) * F;
    const float* base = data + r0 * F;
    for (long long e = threadIdx.x; e < e1; e += active) {
      const double v = (double)__ldg(base + e);
      s += v;
      q += v * v;
    }
    atomicAdd(&sh[col], s);
    atomicAdd(&sh[F + col], q);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) atomicAdd(&sums[i], sh[i]);
}

__global__ void feature_finalize_kernel(const double* __restrict__ sums, long long rows, int F, float* __restrict__ mean,
                                        float* __restrict__ stdv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= F) return;
  const float tot = (float)rows;
  const float m = (float)sums[c] / tot;
  const float var = (float)sums[F + c] / tot - m * m;
  mean[c] = m;
  stdv[c] = sqrtf(fmaxf(var, 0.0f));
}

// ---- pad_collate dict format -> model inputs (SURVEY 8b collate contract iii / 8f-2) ------------------------------
// events[B,T,Fe] (Fe = 14 columns of build_event_features), events_mask[B,T] TRUE = VALID  ->
// x[B,T,7] = selected columns (dt, dt_prev, logflux, logflux_err, band one-hot x3), optional log1p on the two time
// columns, channels 0..3 normalised (x - mean)/(std + 1e-8) on EVERY row (padding included, as the Hyrax collate
// does), and pad[B,T] TRUE = PADDING (the polarity the encoder expects).
__global__ void __launch_bounds__(256) collate_events_kernel(const float* __restrict__ ev, const uint8_t* __restrict__ valid, long long rows,
                                                             int Fe, const int* __restrict__ cols, int log1p_dt,
                                                             const float* __restrict__ mean, const float* __restrict__ stdv,
                                                             float* __restrict__ x, uint8_t* __restrict__ pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 7) return;
  const long long r = i / 7;
  const int c = (int)(i - r * 7);
  float v = ev[r * Fe + cols[c]];
  if (log1p_dt && c < 2) v = log1pf(v);
  if (c < 4) v = (v - mean[c]) / (stdv[c] + 1e-8f);
  x[i] = v;
  if (c == 0) pad[r] = valid[r] ? 0 : 1;
}

}  // namespace

extern "C" {

int acb_prep_lightcurve(const float* raw, const long long* offsets, int B, float horizon, const float* mean, const float* stdv,
                        int max_len, float* x, uint8_t* mask, int* lengths, void* stream) {
  ACB_CHECK(raw && offsets && mean && stdv && x && mask && B > 0 && max_len > 0, "acb_prep_lightcurve: bad arguments");
  prep_lightcurve_kernel<<<cdiv(B, 8), 256, 0, (cudaStream_t)stream>>>(raw, offsets, B, horizon, mean, stdv, max_len, x, mask, lengths);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_prep_events(const double* mjd, const double* mag, const double* magerr, const int* fid, const long long* offsets, int B,
                    long long total, double dt_days, double* tmp, signed char* tmp_b, float* dt, float* dt_prev,
                    signed char* band_id, float* logflux, float* logflux_err, int* n_events, void* stream) {
  ACB_CHECK(mjd && mag && magerr && fid && offsets && tmp && tmp_b && dt && dt_prev && band_id && logflux && logflux_err && n_events && B > 0,
            "acb_prep_events: bad arguments");
  prep_events_kernel<<<cdiv(B, P2_WARPS), 32 * P2_WARPS, 0, (cudaStream_t)stream>>>(mjd, mag, magerr, fid, offsets, B, dt_days, tmp, tmp_b, total, dt,
                                                                      dt_prev, band_id, logflux, logflux_err, n_events);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_prep_spectrum_resample(const double* wl, const double* fx, const long long* offsets, int B, int max_n, const float* grid,
                               int n_grid, float* out, void* stream) {
  return acb_prep_spectrum_resample_idx(wl, fx, offsets, B, max_n, grid, n_grid, out, nullptr, stream);
}

int acb_prep_spectrum_resample_idx(const double* wl, const double* fx, const long long* offsets, int B, int max_n, const float* grid,
                                   int n_grid, float* out, int* idx_out, void* stream) {
  ACB_CHECK(wl && fx && offsets && grid && out && B > 0 && n_grid > 0 && max_n > 0, "acb_prep_spectrum_resample: bad arguments");
  int cap = 64;
  while (cap < max_n) cap <<= 1;
  const size_t smem = (size_t)(2 * cap + n_grid + 34) * 8 + (256 + 8 + 16) * 4 + 258 * 8;
  ACB_CHECK(smem <= 220 * 1024, "acb_prep_spectrum_resample: spectrum too long for shared memory (max_n=%d, n_grid=%d)", max_n, n_grid);
  if (n_grid <= SP_PER * 256) {
    const int cap_r = (max_n + 1) & ~1;  // (its sorting network takes any length: no power-of-two padding)
    const size_t smem_r = std::max((size_t)2 * cap_r * 8, (size_t)n_grid * 4);  // xs | ys, reused as the float staging row
    auto kr = prep_spectrum_reg_kernel;
    ACB_CUDA(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    kr<<<B, 256, smem_r, (cudaStream_t)stream>>>(wl, fx, offsets, cap_r, grid, n_grid, out, idx_out);
    ACB_LAUNCH_CHECK();
    acb_count_launch();
    return ACB_OK;
  }
  auto k = prep_spectrum_kernel;
  ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<B, 256, smem, (cudaStream_t)stream>>>(wl, fx, offsets, cap, grid, n_grid, out, idx_out);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_prep_cutout_norm(const float* img, int B, int C, int H, int W, int cutout_size, int mode, float* out, void* stream) {
  ACB_CHECK(img && out && B > 0 && C > 0 && H == W && H > 0, "acb_prep_cutout_norm: bad arguments");
  ACB_CHECK(mode >= 0 && mode <= 2, "acb_prep_cutout_norm: mode must be 0 (median/std), 1 (L2) or 2 (notebook median/std)");
  int i1 = 0, i2 = H;
  if (cutout_size != H) {
    i1 = (int)((H - cutout_size) / 2.0);  // int((63 - cutout_size) / 2) in the reference (float division, truncation)
    i2 = H - i1;
  }
  const int S = i2 - i1;
  ACB_CHECK(S > 0 && i1 >= 0, "acb_prep_cutout_norm: bad crop");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 1) {
    cutout_l2_kernel<<<B, 256, 0, st>>>(img, C, H, W, i1, S, out);
  } else {
    if (S * S <= CUT_EPT * 256) {
      const int planes = B * C;
      cutout_median_reg_kernel<<<planes < 148 * 5 ? planes : 148 * 5, 256, 0, st>>>(img, planes, H, W, i1, S, mode, out);
      ACB_LAUNCH_CHECK();
      acb_count_launch();
      return ACB_OK;
    }
    const size_t smem = (size_t)((S * S + 3) & ~3) * 4 + 34 * 8 + (256 + 8 + 258) * 4;
    auto k = cutout_median_kernel;
    if (smem > 48 * 1024) ACB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<B * C, 256, smem, st>>>(img, C, H, W, i1, S, mode, out);
  }
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

int acb_feature_stats(const float* data, long long rows, int F, double* work, float* mean, float* stdv, void* stream) {
  ACB_CHECK(data && work && mean && stdv && rows > 0 && F > 0 && F <= 256, "acb_feature_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ACB_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * 2 * F, st));
  const int rows_per_block = 2048;
  feature_sums_kernel<<<cdiv(rows, rows_per_block), 256, 0, st>>>(data, rows, F, rows_per_block, work);
  ACB_LAUNCH_CHECK();
  feature_finalize_kernel<<<cdiv(F, 64), 64, 0, st>>>(work, rows, F, mean, stdv);
  ACB_LAUNCH_CHECK();
  acb_count_launch(2);
  return ACB_OK;
}

int acb_collate_events(const float* events, const uint8_t* valid_mask, int B, int T, int Fe, const int* cols, int log1p_dt,
                       const float* mean, const float* stdv, float* x, uint8_t* pad, void* stream) {
  ACB_CHECK(events && valid_mask && cols && mean && stdv && x && pad && B > 0 && T > 0 && Fe > 0, "acb_collate_events: bad arguments");
  const long long rows = (long long)B * T;
  collate_events_kernel<<<cdiv(rows * 7, 256), 256, 0, (cudaStream_t)stream>>>(events, valid_mask, rows, Fe, cols, log1p_dt, mean, stdv, x, pad);
  ACB_LAUNCH_CHECK();
  acb_count_launch();
  return ACB_OK;
}

}  // extern "C"
