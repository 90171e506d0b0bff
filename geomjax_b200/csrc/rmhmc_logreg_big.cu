// rmhmc for Bayesian logistic regression (Fisher metric) at 32 < D <= 124 -- the shape of BASELINE.json's
// c5 (D = 100, N = 10,000).  Same semantics and parity target as rmhmc_logreg.cu
// (rmhmc/rmhmc.py:131-174, rmhmc/integrators.py:53-156, rmhmc/metrics.py:42-129 on the NEW target of
// SURVEY Appendix B.1); what changes is where the data lives:
//   * the design matrix (4 MB at c5) does not fit shared memory: it is streamed from L2 in tiles of 128 data
//     rows, twice per evaluation of the fixed-point map (pass A: eta, gradient, metric; pass B: h_n, dT/dq);
//   * per chain (= per CTA) shared memory holds s_n = sigmoid(eta_n) for all N rows, the D x D metric (which
//     becomes its Cholesky factor and then G^-1 in place), one X tile (which doubles as the buffer for L^-1);
//   * G = sum_n w_n x_n x_n^T runs as a SYRK on the tile scaled by sqrt(w_n), 4x4 register tiles with rows
//     interleaved (i = bi + NBK a) so that the float4 loads along the data-row axis are bank-conflict free;
//   * h_n = x_n^T G^-1 x_n runs as (tile x G^-1) in 8 x 4 register tiles followed by a row-wise dot product;
//   * Cholesky, triangular inverse and G^-1 = L^-T L^-1 are CTA-wide shared-memory routines.
// FP32 pipe only: the per-chain path for launches without a lock-step plan (small batches, tests); the product path
// for many chains is rmhmc_lockstep.cu (both D^2 N products on the tcgen05 GEMMs of fisher_tc.cu).
#include "launch.h"

namespace gb {

constexpr int BG_THREADS = 512;
constexpr int BG_TN = 128;   // data rows per tile
constexpr int BG_LD = 132;   // row stride (floats) of the tile and metric buffers: 16-byte aligned rows
constexpr int BG_ROWS = 128; // rows of both buffers
constexpr int BG_DMAX = 124;
constexpr int BG_NV = 18;    // small vectors of 128 floats

struct BigLR {
  const float* Xt;  // [D, ldx]
  const float* y;   // [N]
  int N, D, ldx, NBK, ntiles;
  float alpha;
};

struct BGSmem {
  float *sv;   // [N4] sigmoid(eta_n)
  float *G;    // [128][BG_LD] metric -> Cholesky factor (lower) -> inverse metric
  float *T;    // [128][BG_LD] X tile (feature-major: T[i][c] = X[n0 + c, i]); L^-1 between the passes
  float *vec;  // BG_NV vectors of 128
  unsigned char *tbi, *tbj;  // SYRK tile -> (bi, bj)
  __device__ float* v(int i) const { return vec + i * 128; }
};
enum { B_Q = 0, B_P, B_Q0, B_P0, B_QN, B_PN, B_G, B_W, B_DT, B_Z, B_RT, B_SW, B_HB, B_COL, B_PART /* 4 vectors */ };

__device__ __forceinline__ float bg_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < BG_THREADS / 32) ? red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

// T[i][c] = X[n0 + c, i] for i < D, zero elsewhere (rows D..127, columns past N)
__device__ __forceinline__ void bg_load_tile(const BigLR& tg, const BGSmem& sm, int n0) {
  for (int e = threadIdx.x; e < BG_ROWS * (BG_TN / 4); e += BG_THREADS) {
    const int i = e >> 5, c4 = e & 31;
    const int n = n0 + 4 * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < tg.D && n < tg.ldx) {
      v = __ldg((const float4*)(tg.Xt + (size_t)i * tg.ldx + n));
      if (n + 0 >= tg.N) v.x = 0.f;
      if (n + 1 >= tg.N) v.y = 0.f;
      if (n + 2 >= tg.N) v.z = 0.f;
      if (n + 3 >= tg.N) v.w = 0.f;
    }
    *(float4*)(sm.T + i * BG_LD + 4 * c4) = v;
  }
}

// out[c] = sum_i T[i][c] * vec[i] for the 128 data rows of the tile, all 512 threads (4 interleaved
// partial sums per row, combined through shared memory).  Ends with the partials published: the caller
// reads bg_tile_dot_result(sm, c) after the __syncthreads() inside.
__device__ __forceinline__ void bg_tile_dot(const BigLR& tg, const BGSmem& sm, const float* vec) {
  const int c = threadIdx.x & (BG_TN - 1), part = threadIdx.x >> 7;
  float a = 0.f;
  for (int i = part; i < tg.D; i += 4) a = fmaf(sm.T[i * BG_LD + c], vec[i], a);
  sm.v(B_PART)[part * 128 + c] = a;
  __syncthreads();
}
__device__ __forceinline__ float bg_tile_dot_result(const BGSmem& sm, int c) {
  const float* pp = sm.v(B_PART);
  return (pp[c] + pp[128 + c]) + (pp[256 + c] + pp[384 + c]);
}

// sum over the tile's rows of T[i][c] * s[c] for the features owned by this warp (i = warp + 16 k), added to acc[k]
__device__ __forceinline__ void bg_feature_dots(const BigLR& tg, const BGSmem& sm, const float* s, float (&acc)[8]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 s4 = *(const float4*)(s + 4 * lane);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = warp + 16 * k;
    if (i < tg.D) {
      const float4 x4 = *(const float4*)(sm.T + i * BG_LD + 4 * lane);
      float a = x4.x * s4.x + x4.y * s4.y + x4.z * s4.z + x4.w * s4.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      acc[k] += a;
    }
  }
}

// pass A: s_n, log-density, gradient and (need_metric) G = X^T diag(w) X + alpha I in sm.G.  Returns logp.
__device__ float bg_pass_a(const BigLR& tg, const BGSmem& sm, const float* qv, float* gout, bool need_metric, float* red) {
  const int tid = threadIdx.x, D = tg.D, N = tg.N;
  float acc[4][4];
#pragma unroll
  for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
    for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] = 0.f;
  float ga[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) ga[k] = 0.f;
  float lp = 0.f;
  const bool my_tile = need_metric && tid < tg.ntiles;
  const int bi = sm.tbi[tid], bj = sm.tbj[tid];
  float* rt = sm.v(B_RT);
  float* sw = sm.v(B_SW);
  for (int n0 = 0; n0 < N; n0 += BG_TN) {
    bg_load_tile(tg, sm, n0);
    __syncthreads();
    bg_tile_dot(tg, sm, qv);
    if (tid < BG_TN) {
      const int n = n0 + tid;
      const float eta = bg_tile_dot_result(sm, tid);
      float r = 0.f, s_w = 0.f;
      if (n < N) {
        const float yn = __ldg(tg.y + n);
        lp += yn * eta - (fmaxf(eta, 0.f) + log1pf(expf(-fabsf(eta))));  // jnp.logaddexp(0, eta)
        const float s = 1.f / (1.f + expf(-eta));
        sm.sv[n] = s;
        r = yn - s;
        s_w = sqrtf(s * (1.f - s));
      }
      rt[tid] = r;
      sw[tid] = s_w;
    }
    __syncthreads();
    bg_feature_dots(tg, sm, rt, ga);
    if (need_metric) {
      __syncthreads();
      for (int e = tid; e < D * (BG_TN / 4); e += BG_THREADS) {  // tile *= sqrt(w_n) in place
        const int i = e >> 5, c4 = e & 31;
        float4 v = *(float4*)(sm.T + i * BG_LD + 4 * c4);
        const float4 s4 = *(const float4*)(sw + 4 * c4);
        v.x *= s4.x; v.y *= s4.y; v.z *= s4.z; v.w *= s4.w;
        *(float4*)(sm.T + i * BG_LD + 4 * c4) = v;
      }
      __syncthreads();
      if (my_tile) {
        const float* ri = sm.T + bi * BG_LD;
        const float* rj = sm.T + bj * BG_LD;
        const int rs = tg.NBK * BG_LD;
#pragma unroll 2
        for (int c4 = 0; c4 < BG_TN / 4; ++c4) {
          float4 xi[4], xj[4];
#pragma unroll
          for (int a_ = 0; a_ < 4; ++a_) {
            xi[a_] = *(const float4*)(ri + a_ * rs + 4 * c4);
            xj[a_] = *(const float4*)(rj + a_ * rs + 4 * c4);
          }
#pragma unroll
          for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < 4; ++b_) {
              float s = acc[a_][b_];
              s = fmaf(xi[a_].x, xj[b_].x, s);
              s = fmaf(xi[a_].y, xj[b_].y, s);
              s = fmaf(xi[a_].z, xj[b_].z, s);
              s = fmaf(xi[a_].w, xj[b_].w, s);
              acc[a_][b_] = s;
            }
        }
      }
    }
    __syncthreads();
  }
  lp = bg_block_sum(lp, red);
  float qq = 0.f;
  for (int i = 0; i < D; ++i) qq = fmaf(qv[i], qv[i], qq);
  lp -= 0.5f * tg.alpha * qq;
  {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = warp + 16 * k;
      if (lane == 0 && i < D) gout[i] = ga[k] - tg.alpha * qv[i];
    }
  }
  if (my_tile) {
#pragma unroll
    for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < 4; ++b_) {
        const int i = bi + tg.NBK * a_, j = bj + tg.NBK * b_;
        if (bi != bj || i <= j) {  // diagonal tile blocks compute both (i, j) and (j, i): keep one
          const float v = acc[a_][b_] + (i == j ? tg.alpha : 0.f);
          sm.G[i * BG_LD + j] = v;
          sm.G[j * BG_LD + i] = v;
        }
      }
  }
  __syncthreads();
  return lp;
}

// in-place lower Cholesky of sm.G[0:D, 0:D] by the whole CTA; returns log det G (NaN for a non-SPD matrix,
// like jnp.linalg.cholesky)
__device__ float bg_cholesky(const BigLR& tg, const BGSmem& sm, float* red) {
  const int D = tg.D, tid = threadIdx.x;
  float* col = sm.v(B_COL);
  const int warp = tid >> 5, lane = tid & 31;
  for (int k = 0; k < D; ++k) {
    const float lkk = sqrtf(sm.G[k * BG_LD + k]);
    const float rk = 1.f / lkk;
    __syncthreads();
    if (tid == 0) sm.G[k * BG_LD + k] = lkk;
    for (int i = k + 1 + tid; i < D; i += BG_THREADS) {
      const float l = sm.G[i * BG_LD + k] * rk;
      sm.G[i * BG_LD + k] = l;
      col[i] = l;  // contiguous copy of column k: the trailing update reads it conflict-free
    }
    __syncthreads();
    for (int i = k + 1 + warp; i < D; i += BG_THREADS / 32) {  // warp = row, lanes = columns k+1..i
      const float lik = col[i];
      for (int j = k + 1 + lane; j <= i; j += 32) sm.G[i * BG_LD + j] = fmaf(-lik, col[j], sm.G[i * BG_LD + j]);
    }
    __syncthreads();
  }
  float ld = 0.f;
  for (int i = tid; i < D; i += BG_THREADS) ld += logf(sm.G[i * BG_LD + i]);
  return 2.f * bg_block_sum(ld, red);
}

// L (lower, in sm.G) -> L^-1 in sm.T -> G^-1 = L^-T L^-1 in sm.G, zero outside [0:D, 0:D]
__device__ void bg_inverse(const BigLR& tg, const BGSmem& sm) {
  const int D = tg.D, tid = threadIdx.x;
  if (tid < D) {  // column tid of L^-1 by forward substitution
    const int c = tid;
    for (int i = c; i < D; ++i) {
      float s = (i == c) ? 1.f : 0.f;
      for (int k = c; k < i; ++k) s = fmaf(-sm.G[i * BG_LD + k], sm.T[k * BG_LD + c], s);
      sm.T[i * BG_LD + c] = s / sm.G[i * BG_LD + i];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < BG_ROWS * BG_LD; idx += BG_THREADS) {
    const int i = idx / BG_LD, j = idx - i * BG_LD;
    if (i >= D || j >= D) sm.G[idx] = 0.f;
  }
  for (int idx = tid; idx < D * D; idx += BG_THREADS) {
    const int i = idx / D, j = idx - i * D;
    if (j <= i) {
      float s = 0.f;
      for (int k = i; k < D; ++k) s = fmaf(sm.T[k * BG_LD + i], sm.T[k * BG_LD + j], s);
      sm.G[i * BG_LD + j] = s;
      sm.G[j * BG_LD + i] = s;
    }
  }
  __syncthreads();
}

__device__ void bg_matvec(const BigLR& tg, const BGSmem& sm, const float* p, float* w) {
  if (threadIdx.x < tg.D) {
    float s = 0.f;
    for (int j = 0; j < tg.D; ++j) s = fmaf(sm.G[threadIdx.x * BG_LD + j], p[j], s);
    w[threadIdx.x] = s;
  }
  __syncthreads();
}

// pass B: dT/dq_i = 1/2 sum_n w'_n x_ni (h_n - u_n^2), u = X w, h_n = x_n^T G^-1 x_n (sm.G = G^-1, sm.sv = s_n)
__device__ void bg_pass_b(const BigLR& tg, const BGSmem& sm, const float* w, float* dT) {
  const int tid = threadIdx.x, D = tg.D, N = tg.N;
  const int warp = tid >> 5, lane = tid & 31;
  float da[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) da[k] = 0.f;
  float* rt = sm.v(B_RT);
  float* ub = sm.v(B_SW);
  float* hb = sm.v(B_HB);
  for (int n0 = 0; n0 < N; n0 += BG_TN) {
    bg_load_tile(tg, sm, n0);
    if (tid < BG_TN) hb[tid] = 0.f;
    __syncthreads();
    bg_tile_dot(tg, sm, w);
    if (tid < BG_TN) ub[tid] = bg_tile_dot_result(sm, tid);
    {
      // warp = group of 8 metric rows i0..i0+7, lane = 4 consecutive data rows
      const int i0 = 8 * warp;
      if (i0 < D) {
        float4 acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < D; ++j) {
          const float4 x4 = *(const float4*)(sm.T + j * BG_LD + 4 * lane);
          const float4 g0 = *(const float4*)(sm.G + j * BG_LD + i0);      // G^-1 is symmetric: row j, columns i0..
          const float4 g1 = *(const float4*)(sm.G + j * BG_LD + i0 + 4);
          const float gr[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            acc[r].x = fmaf(gr[r], x4.x, acc[r].x);
            acc[r].y = fmaf(gr[r], x4.y, acc[r].y);
            acc[r].z = fmaf(gr[r], x4.z, acc[r].z);
            acc[r].w = fmaf(gr[r], x4.w, acc[r].w);
          }
        }
        float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 x4 = *(const float4*)(sm.T + (i0 + r) * BG_LD + 4 * lane);  // rows >= D are zero
          part.x = fmaf(x4.x, acc[r].x, part.x);
          part.y = fmaf(x4.y, acc[r].y, part.y);
          part.z = fmaf(x4.z, acc[r].z, part.z);
          part.w = fmaf(x4.w, acc[r].w, part.w);
        }
        atomicAdd(hb + 4 * lane + 0, part.x);
        atomicAdd(hb + 4 * lane + 1, part.y);
        atomicAdd(hb + 4 * lane + 2, part.z);
        atomicAdd(hb + 4 * lane + 3, part.w);
      }
    }
    __syncthreads();
    if (tid < BG_TN) {
      const int n = n0 + tid;
      float t = 0.f;
      if (n < N) {
        const float s = sm.sv[n];
        t = s * (1.f - s) * (1.f - 2.f * s) * (hb[tid] - ub[tid] * ub[tid]);
      }
      rt[tid] = t;
    }
    __syncthreads();
    bg_feature_dots(tg, sm, rt, da);
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = warp + 16 * k;
    if (lane == 0 && i < D) dT[i] = 0.5f * da[k];
  }
  __syncthreads();
}

// fixed-point map (rmhmc/integrators.py:119-142): (q, p) -> (qi + he G^-1 p, pi - he (dT/dq - grad))
__device__ void bg_map(const BigLR& tg, const BGSmem& sm, const float* q, const float* p, const float* qi, const float* pi,
                       float he, float* qn, float* pn, float* red) {
  float* g = sm.v(B_G);
  float* w = sm.v(B_W);
  float* dT = sm.v(B_DT);
  bg_pass_a(tg, sm, q, g, true, red);
  bg_cholesky(tg, sm, red);
  bg_inverse(tg, sm);
  bg_matvec(tg, sm, p, w);
  bg_pass_b(tg, sm, w, dT);
  const int tid = threadIdx.x;
  if (tid < tg.D) {
    const float a = qi[tid], b = pi[tid];
    qn[tid] = fmaf(he, w[tid], a);
    pn[tid] = fmaf(-he, dT[tid] - g[tid], b);
  }
  __syncthreads();
}

__device__ float bg_norm(const BigLR& tg, const float* qa, const float* pa, const float* qb, const float* pb, float* red) {
  float mx = 0.f;
  bool nan = false;
  for (int i = threadIdx.x; i < tg.D; i += BG_THREADS) {
    const float dq = fabsf(qa[i] - qb[i]), dp = fabsf(pa[i] - pb[i]);
    nan = nan || isnan(dq) || isnan(dp);
    mx = fmaxf(mx, fmaxf(dq, dp));
  }
  if (nan) mx = __int_as_float(0x7f800000);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = mx;
  __syncthreads();
  float t = red[l & 15];
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  __syncthreads();
  return t;
}

__device__ BGSmem bg_carve(const BigLR& tg, unsigned char* base) {
  BGSmem sm;
  float* f = (float*)base;
  sm.G = f; f += BG_ROWS * BG_LD;
  sm.T = f; f += BG_ROWS * BG_LD;
  sm.vec = f; f += BG_NV * 128;
  sm.sv = f; f += (tg.N + 3) / 4 * 4;
  sm.tbi = (unsigned char*)f;
  sm.tbj = sm.tbi + BG_THREADS;
  return sm;
}
static size_t bg_smem_bytes(int N) {
  return sizeof(float) * (2 * (size_t)BG_ROWS * BG_LD + BG_NV * 128 + (size_t)(N + 3) / 4 * 4) + 2 * BG_THREADS + 64;
}

__device__ void bg_setup_tiles(const BigLR& tg, const BGSmem& sm) {
  const int tid = threadIdx.x;
  int bi = 0, rem = tid;
  if (tid < tg.ntiles) {
    while (rem >= tg.NBK - bi) { rem -= tg.NBK - bi; ++bi; }
    sm.tbi[tid] = (unsigned char)bi;
    sm.tbj[tid] = (unsigned char)(bi + rem);
  } else {
    sm.tbi[tid] = 0;
    sm.tbj[tid] = 0;
  }
  for (int e = tid; e < BG_NV * 128; e += BG_THREADS) sm.vec[e] = 0.f;
  __syncthreads();
}

__global__ void __launch_bounds__(BG_THREADS, 1) rmhmc_logreg_big_kernel(const TransArgs a, const BigLR tg) {
  extern __shared__ __align__(128) unsigned char bg_raw[];
  const BGSmem sm = bg_carve(tg, bg_raw);
  __shared__ float red[32];
  __shared__ float H0_s;
  __shared__ int acc_s;
  bg_setup_tiles(tg, sm);
  const int D = tg.D, tid = threadIdx.x;
  float *q = sm.v(B_Q), *p = sm.v(B_P), *q0 = sm.v(B_Q0), *p0 = sm.v(B_P0), *qn = sm.v(B_QN), *pn = sm.v(B_PN);
  float *g = sm.v(B_G), *w = sm.v(B_W), *z = sm.v(B_Z);
  const float tol = (float)a.fp_tol, div_tol = (float)a.fp_div_tol;
  const long long T = a.ks.keys ? 1 : a.ks.num_transitions;
  // chains are handed out dynamically when the caller gave a workspace (gb200_run_opts.workspace): the
  // fixed-point iteration count is heavy-tailed (a float32 iterate stalling just above tol runs to
  // max_iters), and a static chain -> CTA map leaves most SMs idle behind the slowest CTA
  __shared__ long long chain_s;
  for (long long round = 0;; ++round) {
    if (tid == 0) chain_s = a.work_count ? (long long)atomicAdd(a.work_count, 1) : (long long)blockIdx.x + round * gridDim.x;
    __syncthreads();
    const long long chain = chain_s;
    __syncthreads();
    if (chain >= a.C) break;
    for (long long it = 0; it < T; ++it) {
      const long long t = a.ks.first_transition + it;
      const float* spos = (const float*)(it == 0 ? a.in_pos : a.out_pos) + chain * D;
      const float* slogp = (const float*)(it == 0 ? a.in_logp : a.out_logp);
      const float* sgrad = (const float*)(it == 0 ? a.in_grad : a.out_grad) + chain * D;
      float eps = (float)a.step_size;
      float* da = nullptr;
      if (a.opts.dual_averaging != nullptr) {
        da = (float*)a.opts.dual_averaging + chain * 5;
        eps = expf(da[0]);
      } else if (a.step_size_per_chain != nullptr) {
        eps = ((const float*)a.step_size_per_chain)[chain];
      }
      const float he = 0.5f * eps;
      const float l0 = slogp[chain];
      U2 key = transition_key(a, chain, t);
      U2 k_m, k_a;
      split2(a.mode, key, k_m, k_a);
      if (tid < D) {
        q[tid] = spos[tid];
        if (a.opts.noise_override != nullptr) z[tid] = ((const float*)a.opts.noise_override)[chain * D + tid];
        else z[tid] = bits_to_normal(random_bits_elem(a.mode, k_m, (uint32_t)tid, (uint32_t)D));
        if (a.info.noise) ((float*)a.info.noise)[chain * D + tid] = z[tid];
      }
      __syncthreads();
      // start: metric, momentum p = L z (rmhmc/metrics.py:45-58), H0 = -l0 + T(q, p)
      bg_pass_a(tg, sm, q, g, true, red);
      float logdet = bg_cholesky(tg, sm, red);
      if (tid < D) {
        float s = 0.f;
        for (int j = 0; j <= tid; ++j) s = fmaf(sm.G[tid * BG_LD + j], z[j], s);
        p[tid] = s;
        if (a.info.momentum) ((float*)a.info.momentum)[chain * D + tid] = s;
      }
      __syncthreads();
      bg_inverse(tg, sm);
      bg_matvec(tg, sm, p, w);
      {
        float pw = 0.f;
        for (int i = tid; i < D; i += BG_THREADS) pw = fmaf(p[i], w[i], pw);
        pw = bg_block_sum(pw, red);
        if (tid == 0) H0_s = -l0 + 0.5f * pw + 0.5f * logdet + 0.91893853320467274178f * (float)D;
      }
      __syncthreads();
      int iters_total = 0;
      for (int s = 0; s < a.num_steps; ++s) {  // implicit_midpoint.one_step rmhmc/integrators.py:116-154
        if (tid < D) { q0[tid] = q[tid]; p0[tid] = p[tid]; }
        __syncthreads();
        bg_map(tg, sm, q0, p0, q0, p0, he, q, p, red);
        float nrm = bg_norm(tg, q, p, q0, p0, red);
        int n = 0;
        while ((n < a.fp_max_iters) && (nrm < __int_as_float(0x7f800000)) && (nrm < div_tol) && (nrm > tol)) {
          bg_map(tg, sm, q, p, q0, p0, he, qn, pn, red);
          nrm = bg_norm(tg, qn, pn, q, p, red);
          if (tid < D) { q[tid] = qn[tid]; p[tid] = pn[tid]; }
          __syncthreads();
          ++n;
        }
        iters_total += n;
        bg_map(tg, sm, q, p, q, p, he, qn, pn, red);  // explicit update from the midpoint :147-148
        if (tid < D) { q[tid] = qn[tid]; p[tid] = pn[tid]; }
        __syncthreads();
      }
      // end state: log-density, gradient, velocity, energy, accept
      const float lp = bg_pass_a(tg, sm, q, g, true, red);
      logdet = bg_cholesky(tg, sm, red);
      bg_inverse(tg, sm);
      bg_matvec(tg, sm, p, w);
      float pw = 0.f;
      for (int i = tid; i < D; i += BG_THREADS) pw = fmaf(p[i], w[i], pw);
      pw = bg_block_sum(pw, red);
      if (tid == 0) {
        const float H1 = -lp + 0.5f * pw + 0.5f * logdet + 0.91893853320467274178f * (float)D;
        MH<float> mh = metropolis<float>(a, k_a, chain, H0_s, H1);
        acc_s = mh.accept;
        store_scalar<float>(a.info.acceptance_rate, chain, mh.p_accept);
        if (a.info.is_accepted) a.info.is_accepted[chain] = mh.accept;
        if (a.info.is_divergent) a.info.is_divergent[chain] = mh.divergent;
        store_scalar<float>(a.info.energy, chain, H1);
        store_scalar<float>(a.info.proposal_logdensity, chain, lp);
        store_scalar<float>(a.info.proposal_weight, chain, mh.weight);
        store_scalar<float>(a.info.initial_energy, chain, H0_s);
        store_scalar<float>(a.info.accept_uniform, chain, mh.u);
        if (a.info.fp_iters) a.info.fp_iters[chain] = iters_total;
        if (a.opts.sample_accept != nullptr) ((float*)a.opts.sample_accept)[it * a.C + chain] = mh.p_accept;
        if (a.opts.accept_sum != nullptr) ((float*)a.opts.accept_sum)[chain] += mh.p_accept;
        if (da != nullptr)
          dual_averaging_update<float>(da, mh.p_accept, (float)a.opts.da_target, (float)a.opts.da_t0,
                                       (float)a.opts.da_gamma, (float)a.opts.da_kappa);
        ((float*)a.out_logp)[chain] = mh.accept ? lp : l0;
      }
      __syncthreads();
      if (tid < D) {
        if (a.info.proposal_position) ((float*)a.info.proposal_position)[chain * D + tid] = q[tid];
        if (a.info.proposal_momentum) ((float*)a.info.proposal_momentum)[chain * D + tid] = -p[tid];
        if (a.info.proposal_velocity) ((float*)a.info.proposal_velocity)[chain * D + tid] = -w[tid];
        if (a.info.proposal_logdensity_grad) ((float*)a.info.proposal_logdensity_grad)[chain * D + tid] = g[tid];
        const float qo = acc_s ? q[tid] : spos[tid];
        const float go = acc_s ? g[tid] : sgrad[tid];
        ((float*)a.out_pos)[chain * D + tid] = qo;
        ((float*)a.out_grad)[chain * D + tid] = go;
        if (a.opts.samples != nullptr) ((float*)a.opts.samples)[(it * a.C + chain) * D + tid] = qo;
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(BG_THREADS, 1) init_logreg_big_kernel(const BigLR tg, gb200_state st, long long C) {
  extern __shared__ __align__(128) unsigned char bg_raw[];
  const BGSmem sm = bg_carve(tg, bg_raw);
  __shared__ float red[32];
  bg_setup_tiles(tg, sm);
  float *q = sm.v(B_Q), *g = sm.v(B_G);
  for (long long chain = blockIdx.x; chain < C; chain += gridDim.x) {
    if (threadIdx.x < tg.D) q[threadIdx.x] = ((const float*)st.position)[chain * tg.D + threadIdx.x];
    __syncthreads();
    const float lp = bg_pass_a(tg, sm, q, g, false, red);
    if (threadIdx.x < tg.D) ((float*)st.logdensity_grad)[chain * tg.D + threadIdx.x] = g[threadIdx.x];
    if (threadIdx.x == 0) {
      ((float*)st.logdensity)[chain] = lp;
      if (st.volume_adjustment) ((float*)st.volume_adjustment)[chain] = 0.f;
    }
    __syncthreads();
  }
}

static int bg_setup(const gb200_target_desc& t, BigLR* tg, size_t* smem) {
  if (!t.vec0 || !t.y || t.N < 1) { set_error("logreg: needs vec0 = X^T [D, ldx] (ldx = params[1]) and y [N]"); return GB200_ERR_INVALID_ARGUMENT; }
  if (t.D > BG_DMAX) { set_error("logreg: D=%d > %d is not built in this version", t.D, BG_DMAX); return GB200_ERR_UNSUPPORTED; }
  tg->Xt = (const float*)t.vec0;
  tg->y = (const float*)t.y;
  tg->N = (int)t.N;
  tg->D = t.D;
  tg->ldx = (int)t.params[1];
  tg->alpha = (float)t.params[0];
  tg->NBK = (t.D + 3) / 4;
  tg->ntiles = tg->NBK * (tg->NBK + 1) / 2;
  if (tg->ldx < tg->N || tg->ldx % 4 != 0) { set_error("logreg: ldx must be >= N and a multiple of 4"); return GB200_ERR_INVALID_ARGUMENT; }
  *smem = bg_smem_bytes(tg->N);
  if (*smem > 227 * 1024) {
    set_error("logreg: N=%d rows do not fit the per-chain shared-memory state of this version", tg->N);
    return GB200_ERR_UNSUPPORTED;
  }
  return GB200_OK;
}

int launch_rmhmc_logreg_big(const TransArgs& a, const gb200_target_desc& t, cudaStream_t s) {
  BigLR tg;
  size_t smem;
  int rc = bg_setup(t, &tg, &smem);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(rmhmc_logreg_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("logreg_big: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  const int grid = (int)(a.C < 148 ? a.C : 148);
  TransArgs c = a;
  c.work_count = nullptr;
  c.work_list = nullptr;
  if (a.opts.workspace != nullptr && a.opts.workspace_bytes >= 16) {
    c.work_count = (int*)a.opts.workspace;
    cudaMemsetAsync(c.work_count, 0, 16, s);
  }
  rmhmc_logreg_big_kernel<<<grid, BG_THREADS, smem, s>>>(c, tg);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

int launch_init_logreg_big(const gb200_target_desc& t, gb200_state st, long long C, cudaStream_t s) {
  BigLR tg;
  size_t smem;
  int rc = bg_setup(t, &tg, &smem);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(init_logreg_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("logreg_big: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  const int grid = (int)(C < 148 ? C : 148);
  init_logreg_big_kernel<<<grid, BG_THREADS, smem, s>>>(tg, st, C);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

}  // namespace gb
