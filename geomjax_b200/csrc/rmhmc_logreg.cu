// rmhmc for Bayesian logistic regression with the Fisher-information(+prior) metric
//   G(theta) = X^T diag(s(1-s)) X + alpha I          (NEW target)
// Dense per-chain metric: ONE CTA PER CHAIN, persistent over chains; the shared design matrix is
// staged once per CTA into shared memory with bulk-TMA copies (cp.async.bulk -> UBLKCP) and stays
// resident; the chain's D x D metric, its Cholesky factor, triangular inverse, log-det and the
// trace / quadratic-form terms of dT/dq live in shared memory for the whole trajectory.
// Semantics: rmhmc/rmhmc.py:131-174, rmhmc/integrators.py:53-156, rmhmc/metrics.py:42-129 with
//   dT/dp = G^-1 p = v,   dT/dq_i = 1/2 sum_n w'_n x_ni (h_n - u_n^2),
//   u = X v,  h_n = x_n^T G^-1 x_n = |L^-1 x_n|^2,  w' = s(1-s)(1-2s).
// This round: FP32 CUDA cores (parity first).  The two D^2 N products per evaluation (SYRK for G
// and the h_n quadratic forms) are the tcgen05 candidates named by the north star.
#include "launch.h"

namespace gb {

struct LogRegDev {
  const float* Xt;   // [D, ldx] transposed design matrix (row i = feature i over the N data rows)
  const float* y;    // [N]
  int N, D, ldx;
  float alpha;
};

constexpr int LR_BLOCK = 256;
constexpr int LR_DMAX = 32;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = (l < LR_BLOCK / 32) ? scratch[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// Shared-memory view of one CTA
struct LRSmem {
  float *Xs, *wv, *wpv, *rv, *G, *Li, *vec, *scratch;
  int ldg;
  // vec layout: 12 vectors of LR_DMAX
  __device__ float* v(int i) const { return vec + i * LR_DMAX; }
};
enum { V_Q = 0, V_P, V_Q0, V_P0, V_QN, V_PN, V_G, V_W, V_DT, V_Z, V_TMP, V_RINV, V_NUM };

// logdensity, gradient and the metric pieces at position qv:
//   fills wv (w), wpv (w'), rv (y - s); returns logp; grad -> gout; if need_metric: G, Cholesky L (in G),
//   L^-1 (Li), logdet.
template <int DP>
__device__ float lr_eval(const LogRegDev& tg, const LRSmem& sm, const float* qv, float* gout, bool need_metric,
                         float* logdet_out) {
  const int N = tg.N, D = tg.D, ldx = tg.ldx, tid = threadIdx.x;
  float lp = 0.f;
  for (int n = tid; n < N; n += LR_BLOCK) {
    float eta = 0.f;
    for (int i = 0; i < D; ++i) eta = fmaf(sm.Xs[i * ldx + n], qv[i], eta);
    const float yn = tg.y[n];
    // softplus(eta) = max(eta, 0) + log1p(exp(-|eta|))  (= jnp.logaddexp(0, eta))
    const float sp = fmaxf(eta, 0.f) + log1pf(expf(-fabsf(eta)));
    lp += yn * eta - sp;
    const float s = 1.f / (1.f + expf(-eta));
    const float w = s * (1.f - s);
    sm.wv[n] = w;
    sm.wpv[n] = w * (1.f - 2.f * s);
    sm.rv[n] = yn - s;
  }
  lp = block_sum(lp, sm.scratch);  // includes the __syncthreads that publishes wv / rv
  float qq = 0.f;
  for (int i = 0; i < D; ++i) qq = fmaf(qv[i], qv[i], qq);
  lp -= 0.5f * tg.alpha * qq;
  // grad_i = sum_n X[n,i] r_n - alpha q_i : one warp per feature
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < D; i += LR_BLOCK / 32) {
    float a = 0.f;
    for (int n = lane; n < N; n += 32) a = fmaf(sm.Xs[i * ldx + n], sm.rv[n], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) gout[i] = a - tg.alpha * qv[i];
  }
  if (!need_metric) {
    __syncthreads();
    return lp;
  }
  // G_ij = sum_n w_n X[n,i] X[n,j] + alpha delta_ij : 4x4 register tiles, one warp per tile pair
  // work item = (tile pair, half of the data rows): balances 2 * nb(nb+1)/2 items over the 8 warps
  constexpr int nb = DP / 4;
  constexpr int ntiles = nb * (nb + 1) / 2;
  for (int i = tid; i < DP * sm.ldg; i += LR_BLOCK) {
    const int r = i / sm.ldg, c = i - r * sm.ldg;
    sm.G[i] = (r == c && r < DP) ? tg.alpha : 0.f;  // prior precision on the diagonal (also on padded dims)
    sm.Li[i] = 0.f;                                  // staging buffer for the second half of the data rows
  }
  __syncthreads();
  for (int wI = warp; wI < 2 * ntiles; wI += LR_BLOCK / 32) {
    const int tI = wI >> 1, half = wI & 1;
    int ib = 0, rem = tI;
    while (rem >= nb - ib) { rem -= nb - ib; ++ib; }
    const int jb = ib + rem;
    const int n_lo = half ? (N / 2) : 0, n_hi = half ? N : (N / 2);
    float acc[4][4];
#pragma unroll
    for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] = 0.f;
    for (int n = n_lo + lane; n < n_hi; n += 32) {
      const float w = sm.wv[n];
      float xi[4], xj[4];
#pragma unroll
      for (int a_ = 0; a_ < 4; ++a_) {
        const int i = ib * 4 + a_, j = jb * 4 + a_;
        xi[a_] = sm.Xs[i * ldx + n] * w;  // rows D..DP-1 of Xs are zero padding
        xj[a_] = sm.Xs[j * ldx + n];
      }
#pragma unroll
      for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
        for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] = fmaf(xi[a_], xj[b_], acc[a_][b_]);
    }
    // transpose-reduce: 16 values x 32 lanes -> lane pair (2e, 2e+1) holds total of element e
    // (8+4+2+1+1 = 16 shuffles instead of 16 x 5)
    float v16[16];
#pragma unroll
    for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < 4; ++b_) v16[a_ * 4 + b_] = acc[a_][b_];
    float v8[8], v4[4], v2[2];
    {
      const bool up = lane & 16;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float keep = up ? v16[e + 8] : v16[e], send = up ? v16[e] : v16[e + 8];
        v8[e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool up = lane & 8;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float keep = up ? v8[e + 4] : v8[e], send = up ? v8[e] : v8[e + 4];
        v4[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {
      const bool up = lane & 4;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float keep = up ? v4[e + 2] : v4[e], send = up ? v4[e] : v4[e + 2];
        v2[e] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    }
    float v1;
    {
      const bool up = lane & 2;
      const float keep = up ? v2[1] : v2[0], send = up ? v2[0] : v2[1];
      v1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
    // element index held by this lane: bit4 -> +8, bit3 -> +4, bit2 -> +2, bit1 -> +1
    const int e = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    const int i = ib * 4 + (e >> 2), j = jb * 4 + (e & 3);
    float* dst = half ? sm.Li : sm.G;  // the two row halves land in separate buffers (no atomics)
    if ((lane & 1) == 0 && (ib != jb || i <= j)) {
      dst[i * sm.ldg + j] += v1;
      if (i != j) dst[j * sm.ldg + i] += v1;
    }
  }
  __syncthreads();
  for (int i = tid; i < DP * sm.ldg; i += LR_BLOCK) sm.G[i] += sm.Li[i];
  __syncthreads();
  // Cholesky (lower, in place in G's lower triangle), one warp, lane = row
  if (warp == 0) {
    // register-resident right-looking Cholesky: lane = row, the row lives in registers, column k of
    // L is broadcast with shuffles (a shared-memory version spent 38% of the kernel in this serial
    // section: every trailing update was a dependent LDS -> FMA -> STS chain).
    float row[DP];
#pragma unroll
    for (int j = 0; j < DP; ++j) row[j] = (lane < DP) ? sm.G[lane * sm.ldg + j] : 0.f;
    float diag = 1.f;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      const float dkk = __shfl_sync(0xffffffffu, row[k], k);
      const float lkk = sqrtf(dkk);  // NaN for a non-positive pivot, as jnp's cholesky
      const float lik = (lane == k) ? lkk : row[k] * (1.f / lkk);
      row[k] = lik;
      if (lane == k) diag = lkk;
#pragma unroll
      for (int j = k + 1; j < DP; ++j) {
        const float ljk = __shfl_sync(0xffffffffu, lik, j);
        if (lane >= j) row[j] = fmaf(-lik, ljk, row[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < DP; ++j)
      if (lane < DP && j <= lane) sm.G[lane * sm.ldg + j] = row[j];
    if (lane < DP) sm.v(V_RINV)[lane] = 1.f / diag;
    float ld = (lane < D) ? logf(diag) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if (lane == 0) *logdet_out = 2.f * ld;
  }
  __syncthreads();
  return lp;
}

// w = G^-1 p = L^-T (L^-1 p): two warp-parallel triangular solves (lane = row), warp 0 only
__device__ void lr_solve(const LogRegDev& tg, const LRSmem& sm, const float* p, float* w) {
  const int D = tg.D, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const float* rinv = sm.v(V_RINV);
    float acc = (lane < D) ? p[lane] : 0.f;  // forward: L y = p
    float y = 0.f;
    for (int k = 0; k < D; ++k) {
      const float yk = __shfl_sync(0xffffffffu, acc * rinv[k], k);
      if (lane == k) y = yk;
      if (lane > k && lane < D) acc = fmaf(-sm.G[lane * sm.ldg + k], yk, acc);
    }
    acc = y;  // backward: L^T x = y
    float x = 0.f;
    for (int k = D - 1; k >= 0; --k) {
      const float xk = __shfl_sync(0xffffffffu, acc * rinv[k], k);
      if (lane == k) x = xk;
      if (lane < k) acc = fmaf(-sm.G[k * sm.ldg + lane], xk, acc);
    }
    if (lane < D) w[lane] = x;
  }
  __syncthreads();
}

// dT/dq_i = 1/2 sum_n w'_n x_ni (h_n - u_n^2)
template <int DP>
__device__ void lr_dTdq(const LogRegDev& tg, const LRSmem& sm, const float* w, float* dT) {
  const int N = tg.N, D = tg.D, ldx = tg.ldx, tid = threadIdx.x;
  const float* rinv = sm.v(V_RINV);
  // two data rows per thread per pass: every L_ij fetched from shared memory feeds two independent
  // FMA chains (halves the LDS count per FMA and hides the forward-substitution dependency chain)
  for (int n = tid; n < N; n += 2 * LR_BLOCK) {
    const int n2 = n + LR_BLOCK;
    const bool has2 = n2 < N;
    float xa[DP], xb[DP];  // become y = L^-1 x_n in place (padded dims stay 0)
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      xa[i] = (i < D) ? sm.Xs[i * ldx + n] : 0.f;
      xb[i] = (i < D && has2) ? sm.Xs[i * ldx + n2] : 0.f;
    }
    float ua = 0.f, ub = 0.f, ha = 0.f, hb = 0.f;
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      const float wi = (i < D) ? w[i] : 0.f;
      ua = fmaf(xa[i], wi, ua);
      ub = fmaf(xb[i], wi, ub);
    }
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      float sa = xa[i], sb = xb[i];
#pragma unroll
      for (int j = 0; j < i; ++j) {
        const float lij = sm.G[i * sm.ldg + j];
        sa = fmaf(-lij, xa[j], sa);
        sb = fmaf(-lij, xb[j], sb);
      }
      xa[i] = sa * rinv[i];
      xb[i] = sb * rinv[i];
      ha = fmaf(xa[i], xa[i], ha);
      hb = fmaf(xb[i], xb[i], hb);
    }
    sm.rv[n] = sm.wpv[n] * (ha - ua * ua);
    if (has2) sm.rv[n2] = sm.wpv[n2] * (hb - ub * ub);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < D; i += LR_BLOCK / 32) {
    float a = 0.f;
    for (int n = lane; n < N; n += 32) a = fmaf(sm.Xs[i * ldx + n], sm.rv[n], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) dT[i] = 0.5f * a;
  }
  __syncthreads();
}

// fixed-point map (rmhmc/integrators.py:119-142): (q, p) -> (qi + he w, pi - he (dT/dq - grad))
template <int DP>
__device__ void lr_map(const LogRegDev& tg, const LRSmem& sm, const float* q, const float* p, const float* qi,
                       const float* pi, float he, float* qn, float* pn) {
  float ld;
  float* g = sm.v(V_G);
  float* w = sm.v(V_W);
  float* dT = sm.v(V_DT);
  __shared__ float ld_s;
  lr_eval<DP>(tg, sm, q, g, true, &ld_s);
  (void)ld;
  lr_solve(tg, sm, p, w);
  lr_dTdq<DP>(tg, sm, w, dT);
  const int tid = threadIdx.x;
  if (tid < tg.D) {
    const float a = qi[tid], b = pi[tid];  // read before write: qn/pn may alias qi/pi
    qn[tid] = fmaf(he, w[tid], a);
    pn[tid] = fmaf(-he, dT[tid] - g[tid], b);
  }
  __syncthreads();
}

__device__ float lr_norm(const LogRegDev& tg, const LRSmem& sm, const float* qa, const float* pa, const float* qb,
                         const float* pb) {
  __shared__ float nrm_s;
  if (threadIdx.x < 32) {
    float mx = 0.f;
    bool nan = false;
    for (int i = threadIdx.x; i < tg.D; i += 32) {
      const float dq = fabsf(qa[i] - qb[i]), dp = fabsf(pa[i] - pb[i]);
      nan = nan || isnan(dq) || isnan(dp);
      mx = fmaxf(mx, fmaxf(dq, dp));
    }
    if (nan) mx = __int_as_float(0x7f800000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) nrm_s = mx;
  }
  __syncthreads();
  const float r = nrm_s;
  __syncthreads();
  return r;
}

__device__ void lr_stage_X(const LogRegDev& tg, const LRSmem& sm) {
  // bulk-TMA: one cp.async.bulk per feature row, all completing on one mbarrier
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(&mbar);
  const unsigned row_bytes = (unsigned)tg.ldx * 4u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(row_bytes * (unsigned)tg.D) : "memory");
    for (int i = 0; i < tg.D; ++i) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(sm.Xs + (size_t)i * tg.ldx);
      const float* src = tg.Xt + (size_t)i * tg.ldx;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(row_bytes), "r"(mbar_a)
                   : "memory");
    }
  }
  for (int i = tg.D * tg.ldx + threadIdx.x; i < (tg.D + 3) / 4 * 4 * tg.ldx; i += blockDim.x) sm.Xs[i] = 0.f;  // padding rows
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar_a)
        : "memory");
  }
  __syncthreads();
}

__device__ LRSmem lr_carve(const LogRegDev& tg, unsigned char* base) {
  LRSmem sm;
  float* f = (float*)base;
  sm.Xs = f; f += (size_t)((tg.D + 3) / 4 * 4) * tg.ldx;
  sm.wv = f; f += tg.ldx;
  sm.wpv = f; f += tg.ldx;
  sm.rv = f; f += tg.ldx;
  sm.ldg = LR_DMAX + 1;
  sm.G = f; f += LR_DMAX * sm.ldg;
  sm.Li = f; f += LR_DMAX * sm.ldg;
  sm.vec = f; f += V_NUM * LR_DMAX;
  sm.scratch = f;
  return sm;
}

static size_t lr_smem_bytes(int D, int ldx) {
  return sizeof(float) * ((size_t)((D + 3) / 4 * 4) * ldx + 3 * (size_t)ldx + 2 * LR_DMAX * (LR_DMAX + 1) + V_NUM * LR_DMAX + 64);
}

template <int DP>
__global__ void __launch_bounds__(LR_BLOCK, 1) rmhmc_logreg_kernel(const TransArgs a, const LogRegDev tg) {
  extern __shared__ __align__(128) unsigned char lr_smem[];
  const LRSmem sm = lr_carve(tg, lr_smem);
  lr_stage_X(tg, sm);
  const int D = tg.D, tid = threadIdx.x;
  float *q = sm.v(V_Q), *p = sm.v(V_P), *q0 = sm.v(V_Q0), *p0 = sm.v(V_P0), *qn = sm.v(V_QN), *pn = sm.v(V_PN);
  float *g = sm.v(V_G), *w = sm.v(V_W), *z = sm.v(V_Z);
  __shared__ float logdet_s, H0_s, lp_s;
  __shared__ int go_s;
  const float tol = (float)a.fp_tol, div_tol = (float)a.fp_div_tol;
  const long long T = a.ks.keys ? 1 : a.ks.num_transitions;

  const long long nwork = a.work_list ? (long long)*a.work_count : a.C;
  for (long long widx = blockIdx.x; widx < nwork; widx += gridDim.x) {
    const long long chain = a.work_list ? (long long)a.work_list[widx] : widx;
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
      // metric at the start: momentum p = L z (rmhmc/metrics.py:45-58), H0 = -l0 + T(q, p)
      lr_eval<DP>(tg, sm, q, g, true, &logdet_s);
      if (tid < D) {
        float s = 0.f;
        for (int j = 0; j <= tid; ++j) s = fmaf(sm.G[tid * sm.ldg + j], z[j], s);
        p[tid] = s;
        if (a.info.momentum) ((float*)a.info.momentum)[chain * D + tid] = s;
      }
      __syncthreads();
      lr_solve(tg, sm, p, w);
      if (tid == 0) {
        float pw = 0.f;
        for (int i = 0; i < D; ++i) pw = fmaf(p[i], w[i], pw);
        H0_s = -l0 + 0.5f * pw + 0.5f * logdet_s + 0.91893853320467274178f * (float)D;
      }
      __syncthreads();
      int iters_total = 0;
      for (int s = 0; s < a.num_steps; ++s) {
        if (tid < D) { q0[tid] = q[tid]; p0[tid] = p[tid]; }
        __syncthreads();
        lr_map<DP>(tg, sm, q0, p0, q0, p0, he, q, p);
        float nrm = lr_norm(tg, sm, q, p, q0, p0);
        int n = 0;
        while ((n < a.fp_max_iters) && (nrm < __int_as_float(0x7f800000)) && (nrm < div_tol) && (nrm > tol)) {
          lr_map<DP>(tg, sm, q, p, q0, p0, he, qn, pn);
          nrm = lr_norm(tg, sm, qn, pn, q, p);
          if (tid < D) { q[tid] = qn[tid]; p[tid] = pn[tid]; }
          __syncthreads();
          ++n;
        }
        iters_total += n;
        lr_map<DP>(tg, sm, q, p, q, p, he, qn, pn);  // explicit update from the midpoint
        if (tid < D) { q[tid] = qn[tid]; p[tid] = pn[tid]; }
        __syncthreads();
      }
      // end state
      const float lp = lr_eval<DP>(tg, sm, q, g, true, &logdet_s);
      lr_solve(tg, sm, p, w);
      __shared__ float H1_s;
      __shared__ int acc_s;
      if (tid == 0) {
        float pw = 0.f;
        for (int i = 0; i < D; ++i) pw = fmaf(p[i], w[i], pw);
        const float H1 = -lp + 0.5f * pw + 0.5f * logdet_s + 0.91893853320467274178f * (float)D;
        H1_s = H1;
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
        lp_s = lp;
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
  (void)go_s;
}

template <int DP>
__global__ void __launch_bounds__(LR_BLOCK, 1) init_logreg_kernel(const LogRegDev tg, gb200_state st, long long C) {
  extern __shared__ __align__(128) unsigned char lr_smem[];
  const LRSmem sm = lr_carve(tg, lr_smem);
  lr_stage_X(tg, sm);
  float *q = sm.v(V_Q), *g = sm.v(V_G);
  __shared__ float dummy;
  for (long long chain = blockIdx.x; chain < C; chain += gridDim.x) {
    if (threadIdx.x < tg.D) q[threadIdx.x] = ((const float*)st.position)[chain * tg.D + threadIdx.x];
    __syncthreads();
    const float lp = lr_eval<DP>(tg, sm, q, g, false, &dummy);
    if (threadIdx.x < tg.D) ((float*)st.logdensity_grad)[chain * tg.D + threadIdx.x] = g[threadIdx.x];
    if (threadIdx.x == 0) {
      ((float*)st.logdensity)[chain] = lp;
      if (st.volume_adjustment) ((float*)st.volume_adjustment)[chain] = 0.f;
    }
    __syncthreads();
  }
}

static int lr_setup(const gb200_target_desc& t, LogRegDev* tg, size_t* smem) {
  if (!t.vec0 || !t.y || t.N < 1) { set_error("logreg: needs vec0 = X^T [D, ldx] (ldx = params[1]) and y [N]"); return GB200_ERR_INVALID_ARGUMENT; }
  if (t.D > LR_DMAX) { set_error("logreg: D=%d > %d is not built in this version", t.D, LR_DMAX); return GB200_ERR_UNSUPPORTED; }
  tg->Xt = (const float*)t.vec0;
  tg->y = (const float*)t.y;
  tg->N = (int)t.N;
  tg->D = t.D;
  tg->ldx = (int)t.params[1];
  tg->alpha = (float)t.params[0];
  if (tg->ldx < tg->N || tg->ldx % 4 != 0) { set_error("logreg: ldx must be >= N and a multiple of 4"); return GB200_ERR_INVALID_ARGUMENT; }
  *smem = lr_smem_bytes(tg->D, tg->ldx);
  if (*smem > 220 * 1024) {
    set_error("logreg: design matrix (%d x %d) does not fit the shared-memory-resident kernel of this version", tg->N, tg->D);
    return GB200_ERR_UNSUPPORTED;
  }
  return GB200_OK;
}

// rmhmc_logreg_big.cu: D > 32 or a design matrix that does not fit shared memory (X streamed from L2)
int launch_rmhmc_logreg_big(const TransArgs& a, const gb200_target_desc& t, cudaStream_t s);
int launch_init_logreg_big(const gb200_target_desc& t, gb200_state st, long long C, cudaStream_t s);
static bool lr_needs_big(const gb200_target_desc& t) {
  const int ldx = (int)t.params[1];
  return t.D > LR_DMAX || lr_smem_bytes(t.D, ldx) > 220 * 1024;
}

// CTA-per-chain FP32 kernel over all chains (work_list == NULL) or over a device-side work list
static int launch_lr_per_chain(const TransArgs& a, const gb200_target_desc& t, cudaStream_t s) {
  LogRegDev tg;
  size_t smem;
  int rc = lr_setup(t, &tg, &smem);
  if (rc) return rc;
  const int grid = a.work_list ? 148 : (int)(a.C < 148 ? a.C : 148);  // list length is only known on the device
  const int dp = (tg.D + 3) / 4 * 4;
#define GB_LR(DPV)                                                                                              \
  if (dp == DPV) {                                                                                              \
    cudaError_t e = cudaFuncSetAttribute(rmhmc_logreg_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) { set_error("logreg: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }            \
    rmhmc_logreg_kernel<DPV><<<grid, LR_BLOCK, smem, s>>>(a, tg);                                               \
    GB_CHECK_LAUNCH();                                                                                          \
    return GB200_OK;                                                                                            \
  }
  GB_LR(4) GB_LR(8) GB_LR(12) GB_LR(16) GB_LR(20) GB_LR(24) GB_LR(28) GB_LR(32)
#undef GB_LR
  set_error("logreg: unsupported D");
  return GB200_ERR_UNSUPPORTED;
}

int launch_rmhmc_logreg(const TransArgs& a0, const gb200_target_desc& t, int dtype, cudaStream_t s) {
  if (dtype != GB200_F32) { set_error("logreg: float32 only"); return GB200_ERR_UNSUPPORTED; }
  if (t.metric != GB200_METRIC_TARGET) { set_error("logreg: only the Fisher metric is built"); return GB200_ERR_UNSUPPORTED; }
  if (!t.vec0 || !t.y || t.N < 1) { set_error("logreg: needs vec0 = X^T [D, ldx] (ldx = params[1]) and y [N]"); return GB200_ERR_INVALID_ARGUMENT; }
  TransArgs a = a0;
  a.work_count = nullptr;
  a.work_list = nullptr;
  if (lr_needs_big(t)) return launch_rmhmc_logreg_big(a, t, s);
  // CTA-per-chain FP32 kernels: the path of launches WITHOUT a lock-step plan (gb200_run_opts.plan == NULL; small
  // batches, tests).  The product path for many chains is rmhmc_lockstep.cu.
  return launch_lr_per_chain(a, t, s);
}

int launch_init_logreg(const gb200_target_desc& t, gb200_state st, long long C, int dtype, cudaStream_t s) {
  if (dtype != GB200_F32) { set_error("logreg: float32 only"); return GB200_ERR_UNSUPPORTED; }
  if (t.vec0 && t.y && t.N >= 1 && lr_needs_big(t)) return launch_init_logreg_big(t, st, C, s);
  LogRegDev tg;
  size_t smem;
  int rc = lr_setup(t, &tg, &smem);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(init_logreg_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("logreg: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  const int grid = (int)(C < 148 ? C : 148);
  init_logreg_kernel<4><<<grid, LR_BLOCK, smem, s>>>(tg, st, C);  // no metric needed: DP is irrelevant
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

}  // namespace gb
