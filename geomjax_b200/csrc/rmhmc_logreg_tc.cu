// rmhmc for Bayesian logistic regression (Fisher metric) with the two D^2 N products of every
// fixed-point evaluation on the 5th-generation tensor cores.
//
// One CTA owns a TILE of TC_CB = 64 chains and advances them in lock-step (masked commits: exactly
// the semantics of a vmapped while_loop, rmhmc/integrators.py:53-89).  Per evaluation of the
// implicit-midpoint map (rmhmc/integrators.py:119-142) the chain-shared operand structure is
//     (1) vec(G)[pair, chain] = sum_n Z[n, pair] w[n, chain]            Z[n,(i,j)] = x_ni x_nj
//     (2) h[n, chain]         = sum_pair Z[n, pair] m_pair Ginv[pair, chain]   (m = 1 diag, 2 off-diag)
// Both run as tcgen05.mma kind::tf32 (3xTF32: hi*hi + hi*lo + lo*hi) with FP32 accumulators in
// tensor memory; the Khatri-Rao operand Z is built on the fly in shared memory from a staged X
// tile (never materialised), the chain operands (w, Ginv) are produced by the CUDA cores of the
// same CTA.  (1) uses two-level accumulation (TMEM chunk of 128 data rows -> FP32 shared memory)
// because the tensor core accumulates with truncation (see fisher_tc.cu).
// Everything O(N D) or O(D^3) per chain (eta = X theta, gradient, u = X v, dT reduction, Cholesky,
// inverse) stays on the FP32 pipe.  Semantics and parity target are identical to rmhmc_logreg.cu.
#include "launch.h"

namespace gb {

constexpr int TC_CB = 64;         // chains per CTA = MMA N
constexpr int TC_THREADS = 256;
constexpr int TC_KT = 32;         // K elements per stage
constexpr int TC_KC = 4;          // stages per TMEM chunk in (1)
constexpr int TC_DS = 28;         // stride of per-chain D-vectors in shared memory (16-byte aligned rows)
constexpr int TC_LBO = 128;
constexpr int TC_SBO = (TC_KT / 4) * 128;
constexpr int TC_A_BYTES = 128 * TC_KT * 4;    // one A operand tile (hi or lo)
constexpr int TC_B_BYTES = TC_CB * TC_KT * 4;  // one B operand tile
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_CB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct LogRegTC {
  const float* Xt;
  const float* y;
  int N, D, ldx, P, PS;  // P = D(D+1)/2 pairs, PS = P rounded up to a multiple of TC_KT
  float alpha;
};

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((TC_LBO >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((TC_SBO >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ int tc_off(int row, int k) {
  return (row >> 3) * TC_SBO + (k >> 2) * TC_LBO + (row & 7) * 16 + (k & 3) * 4;
}
__device__ __forceinline__ void tc_split(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  lo = __uint_as_float(__float_as_uint(a - hi) & 0xFFFFE000u);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
      : "memory");
}

struct TCSmem {
  unsigned char *A_hi, *A_lo, *B_hi, *B_lo;
  float *xs;      // staged X tile: pass 1 [D][TC_KT], pass 2 [D][128]
  float *Gs;      // [TC_CB][PS] packed symmetric (G, then L is not stored, then Ginv)
  float *q, *p, *q0, *p0, *g, *w, *dT, *z;  // [TC_CB][TC_DS]; z aliases dT (z is consumed before the first pass 2)
  float *lp, *logdet, *nrm, *H0;            // [TC_CB]
  float *red;                               // candidate iterate (qn, pn): [2][TC_CB][TC_DS]
  short *pair_i, *pair_j;                   // [PS]
  uint32_t mbar_a, tmem;
  uint32_t phase;
};

__device__ __forceinline__ int tc_pair_index(int D, int i, int j) {  // i <= j
  return i * D - (i * (i - 1)) / 2 + (j - i);
}

__device__ __forceinline__ void tc_commit_and_wait(TCSmem& sm) {
  if (threadIdx.x == 0)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sm.mbar_a) : "memory");
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(sm.mbar_a), "r"(sm.phase)
        : "memory");
  }
  sm.phase ^= 1u;
}

// ---- pass 1: eta, w, gradient, log-density and vec(G) = Z^T W on the tensor cores -------------------
// Leaves: g[c][i], lp[c], Gs[c][pair] = G (without the prior term).
template <int DP>
__device__ __noinline__ void tc_pass1(const LogRegTC& tg, TCSmem& sm, const float* qv, bool need_metric) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = tg.D, N = tg.N;
  const int c = tid >> 2, kq = tid & 3;  // chain, quarter of the stage's rows (4 adjacent lanes per chain)
  const int MT = (tg.P + 127) / 128;
  float gp[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) gp[i] = 0.f;
  float lp = 0.f;
  float qc[DP];
#pragma unroll
  for (int i = 0; i < DP; ++i) qc[i] = (i < D) ? qv[c * TC_DS + i] : 0.f;
  const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(sm.A_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(sm.A_lo);
  const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(sm.B_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(sm.B_lo);
  const int ktiles = (N + TC_KT - 1) / TC_KT;
  for (int kt = 0; kt < ktiles; ++kt) {
    const int n0 = kt * TC_KT;
    for (int e = tid; e < D * TC_KT; e += TC_THREADS) {
      const int i = e / TC_KT, kk = e - i * TC_KT;
      sm.xs[e] = (n0 + kk < N) ? tg.Xt[(size_t)i * tg.ldx + n0 + kk] : 0.f;
    }
    __syncthreads();
    // chain c, rows kq*8 .. kq*8+7 of this stage
    {
      float eta[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) eta[r] = 0.f;
#pragma unroll
      for (int i = 0; i < DP; ++i) {
        if (i < D) {
          const float4 x0 = *(const float4*)(sm.xs + i * TC_KT + kq * 8);
          const float4 x1 = *(const float4*)(sm.xs + i * TC_KT + kq * 8 + 4);
          eta[0] = fmaf(x0.x, qc[i], eta[0]); eta[1] = fmaf(x0.y, qc[i], eta[1]);
          eta[2] = fmaf(x0.z, qc[i], eta[2]); eta[3] = fmaf(x0.w, qc[i], eta[3]);
          eta[4] = fmaf(x1.x, qc[i], eta[4]); eta[5] = fmaf(x1.y, qc[i], eta[5]);
          eta[6] = fmaf(x1.z, qc[i], eta[6]); eta[7] = fmaf(x1.w, qc[i], eta[7]);
        }
      }
      float rr[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n = n0 + kq * 8 + r;
        float wv = 0.f;
        rr[r] = 0.f;
        if (n < N) {
          const float yn = __ldg(tg.y + n);
          lp += yn * eta[r] - (fmaxf(eta[r], 0.f) + log1pf(expf(-fabsf(eta[r]))));
          const float s = 1.f / (1.f + expf(-eta[r]));
          wv = s * (1.f - s);
          rr[r] = yn - s;
        }
        if (need_metric) {
          float hi, lo;
          tc_split(wv, hi, lo);
          const int off = tc_off(c, kq * 8 + r);
          *(float*)(sm.B_hi + off) = hi;
          *(float*)(sm.B_lo + off) = lo;
        }
      }
#pragma unroll
      for (int i = 0; i < DP; ++i) {
        if (i < D) {
          const float4 x0 = *(const float4*)(sm.xs + i * TC_KT + kq * 8);
          const float4 x1 = *(const float4*)(sm.xs + i * TC_KT + kq * 8 + 4);
          float a = gp[i];
          a = fmaf(x0.x, rr[0], a); a = fmaf(x0.y, rr[1], a); a = fmaf(x0.z, rr[2], a); a = fmaf(x0.w, rr[3], a);
          a = fmaf(x1.x, rr[4], a); a = fmaf(x1.y, rr[5], a); a = fmaf(x1.z, rr[6], a); a = fmaf(x1.w, rr[7], a);
          gp[i] = a;
        }
      }
    }
    if (need_metric) {
      for (int mt = 0; mt < MT; ++mt) {
        // A tile: row = pair mt*128 + (tid & 127); this thread fills 16 of the 32 K columns
        {
          const int row = tid & 127, kh = (tid >> 7) * 16;
          const int pr = mt * 128 + row;
          const bool ok = pr < tg.P;
          const float* xi = sm.xs + (ok ? sm.pair_i[pr] : 0) * TC_KT + kh;
          const float* xj = sm.xs + (ok ? sm.pair_j[pr] : 0) * TC_KT + kh;
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) {
            const float a = ok ? xi[kk] * xj[kk] : 0.f;
            float hi, lo;
            tc_split(a, hi, lo);
            const int off = tc_off(row, kh + kk);
            *(float*)(sm.A_hi + off) = hi;
            *(float*)(sm.A_lo + off) = lo;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t td = sm.tmem + (uint32_t)(mt * TC_CB);
#pragma unroll
          for (int k8 = 0; k8 < TC_KT / 8; ++k8) {
            const uint32_t adv = (uint32_t)k8 * 2u * TC_LBO;
            const uint32_t acc0 = ((kt % TC_KC) > 0 || k8 > 0) ? 1u : 0u;
            tc_mma(td, tc_desc(a_hi_s + adv), tc_desc(b_hi_s + adv), acc0);
            tc_mma(td, tc_desc(a_hi_s + adv), tc_desc(b_lo_s + adv), 1u);
            tc_mma(td, tc_desc(a_lo_s + adv), tc_desc(b_hi_s + adv), 1u);
          }
        }
        tc_commit_and_wait(sm);  // single-buffered operands: the MMAs must finish reading before reuse
      }
      if ((kt % TC_KC) == TC_KC - 1 || kt == ktiles - 1) {
        // drain the chunk into Gs (FP32, round-to-nearest): warp w reads lanes 32(w%4).., columns 32(w/4)..
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool first = kt < TC_KC;
        const int row = (warp & 3) * 32 + lane, cb = (warp >> 2) * 32;
        for (int mt = 0; mt < MT; ++mt) {
          const int pr = mt * 128 + row;
#pragma unroll
          for (int c8 = 0; c8 < 32; c8 += 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(sm.tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * TC_CB + cb + c8)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (pr < tg.P) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float* dst = sm.Gs + (size_t)(cb + c8 + e) * tg.PS + pr;
                *dst = first ? __uint_as_float(r[e]) : *dst + __uint_as_float(r[e]);
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    __syncthreads();
  }
  // reduce the four row-quarters (adjacent lanes): gradient and log-density
  lp += __shfl_xor_sync(0xffffffffu, lp, 1);
  lp += __shfl_xor_sync(0xffffffffu, lp, 2);
  float qq = 0.f;
#pragma unroll
  for (int i = 0; i < DP; ++i) {
    float s = gp[i];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (kq == 0 && i < D) sm.g[c * TC_DS + i] = s - tg.alpha * qc[i];
    qq = fmaf(qc[i], qc[i], qq);
  }
  if (kq == 0) sm.lp[c] = lp - 0.5f * tg.alpha * qq;
  __syncthreads();
}

// ---- per-chain dense algebra (warp per chain, register-resident): Cholesky of G + alpha I, log-det,
// optional momentum draw p = L z, inverse (packed into Gs), w = G^-1 pv ---------------------------------
template <int DP>
__device__ __noinline__ void tc_factor(const LogRegTC& tg, TCSmem& sm, bool draw, const float* pv, float* wout) {
  // Warp per chain.  Column / row broadcasts go through a per-warp shared-memory scratch and are read
  // back as float4 (a first version used one shuffle per matrix element: ~1200 dependent
  // shuffle->FMA pairs per chain, 30% of the kernel).  The scratch lives in the operand-tile region,
  // which is idle during this phase.
  constexpr int LD = DP;                      // row stride of the scratch matrices (DP % 4 == 0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, D = tg.D;
  float* scratch = (float*)sm.A_hi + warp * (DP * LD + 32);
  float* Lm = scratch;                        // L, row-major
  float* XT = scratch;                        // (L^-1)^T: XT[c][k] = X[k][c]; reuses L's storage once X is complete
  float* colb = scratch + DP * LD;            // one column
  for (int c = warp; c < TC_CB; c += TC_THREADS / 32) {
    float* Gc = sm.Gs + (size_t)c * tg.PS;
    float row[DP];
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      float v = 0.f;
      if (lane < D && j <= lane) v = Gc[tc_pair_index(D, j, lane)];
      if (j == lane) v = (lane < D) ? v + tg.alpha : 1.f;  // padded dims: identity
      row[j] = v;
    }
    float diag = 1.f;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      const float dkk = __shfl_sync(0xffffffffu, row[k], k);
      const float lkk = sqrtf(dkk);
      const float lik = (lane == k) ? lkk : row[k] * (1.f / lkk);
      row[k] = lik;
      if (lane == k) diag = lkk;
      if (lane < DP) colb[lane] = lik;
      __syncwarp();
#pragma unroll
      for (int j4 = (k + 1) / 4; j4 < DP / 4; ++j4) {
        const float4 cj = *(const float4*)(colb + 4 * j4);
        if (4 * j4 + 0 > k && lane >= 4 * j4 + 0) row[4 * j4 + 0] = fmaf(-lik, cj.x, row[4 * j4 + 0]);
        if (4 * j4 + 1 > k && lane >= 4 * j4 + 1) row[4 * j4 + 1] = fmaf(-lik, cj.y, row[4 * j4 + 1]);
        if (4 * j4 + 2 > k && lane >= 4 * j4 + 2) row[4 * j4 + 2] = fmaf(-lik, cj.z, row[4 * j4 + 2]);
        if (4 * j4 + 3 > k && lane >= 4 * j4 + 3) row[4 * j4 + 3] = fmaf(-lik, cj.w, row[4 * j4 + 3]);
      }
      __syncwarp();
    }
    float ld = (lane < D) ? logf(diag) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if (lane == 0) sm.logdet[c] = 2.f * ld;
    if (draw) {  // p = L z (rmhmc/metrics.py:45-58)
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < DP; ++j) s = fmaf((j <= lane) ? row[j] : 0.f, (j < D) ? sm.z[c * TC_DS + j] : 0.f, s);
      if (lane < D) sm.p[c * TC_DS + lane] = s;
    }
    // L (lower, zeros above the diagonal) to the scratch, row-major
    if (lane < DP) {
#pragma unroll
      for (int j4 = 0; j4 < DP / 4; ++j4) {
        float4 v;
        v.x = (4 * j4 + 0 <= lane) ? row[4 * j4 + 0] : 0.f;
        v.y = (4 * j4 + 1 <= lane) ? row[4 * j4 + 1] : 0.f;
        v.z = (4 * j4 + 2 <= lane) ? row[4 * j4 + 2] : 0.f;
        v.w = (4 * j4 + 3 <= lane) ? row[4 * j4 + 3] : 0.f;
        *(float4*)(Lm + lane * LD + 4 * j4) = v;
      }
    }
    __syncwarp();
    // X = L^-1 by forward substitution: lane = column, X[i] = entry (i, lane); row i of L is a broadcast read
    float X[DP];
    const float rdiag = 1.f / diag;
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      float s0 = (i == lane) ? 1.f : 0.f, s1 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < (i + 3) / 4; ++k4) {
        const float4 l4 = *(const float4*)(Lm + i * LD + 4 * k4);
        if (4 * k4 + 0 < i) s0 = fmaf(-l4.x, X[4 * k4 + 0], s0);
        if (4 * k4 + 1 < i) s1 = fmaf(-l4.y, X[4 * k4 + 1], s1);
        if (4 * k4 + 2 < i) s0 = fmaf(-l4.z, X[4 * k4 + 2], s0);
        if (4 * k4 + 3 < i) s1 = fmaf(-l4.w, X[4 * k4 + 3], s1);
      }
      const float ri = __shfl_sync(0xffffffffu, rdiag, i);
      X[i] = (lane <= i) ? (s0 + s1) * ri : 0.f;
    }
    // (L^-1)^T to the scratch: row c = column c of X
    __syncwarp();
    if (lane < DP) {
#pragma unroll
      for (int k4 = 0; k4 < DP / 4; ++k4)
        *(float4*)(XT + lane * LD + 4 * k4) = make_float4(X[4 * k4], X[4 * k4 + 1], X[4 * k4 + 2], X[4 * k4 + 3]);
    }
    __syncwarp();
    // Ginv[i][lane] = sum_k X[k][i] X[k][lane]; column i of X = row i of XT (broadcast float4 reads)
#pragma unroll
    for (int i = 0; i < DP; ++i) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k4 = i / 4; k4 < DP / 4; ++k4) {
        const float4 x4 = *(const float4*)(XT + i * LD + 4 * k4);
        a0 = fmaf(x4.x, X[4 * k4 + 0], a0);
        a1 = fmaf(x4.y, X[4 * k4 + 1], a1);
        a0 = fmaf(x4.z, X[4 * k4 + 2], a0);
        a1 = fmaf(x4.w, X[4 * k4 + 3], a1);
      }
      if (lane < D && i <= lane) Gc[tc_pair_index(D, i, lane)] = a0 + a1;
    }
    __syncwarp();
    if (pv != nullptr) {  // w = Ginv pv
      float s = 0.f;
      if (lane < D) {
        for (int i = 0; i < D; ++i) {
          const int lo = i < lane ? i : lane, hi = i < lane ? lane : i;
          s = fmaf(Gc[tc_pair_index(D, lo, hi)], (draw ? sm.p : pv)[c * TC_DS + i], s);
        }
        wout[c * TC_DS + lane] = s;
      }
    }
    __syncwarp();
  }
  __syncthreads();
}

// ---- pass 2: h = Z (m . Ginv) on the tensor cores, then dT/dq_i = 1/2 sum_n w'_n x_ni (h_n - u_n^2) --
template <int DP>
__device__ __noinline__ void tc_pass2(const LogRegTC& tg, TCSmem& sm, const float* qv) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = tg.D, N = tg.N;
  const int NT = (N + 127) / 128, KP = tg.PS / TC_KT;
  const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(sm.A_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(sm.A_lo);
  const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(sm.B_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(sm.B_lo);
  // data-row tile outermost: the X tile is staged from L2 once per 128 rows (it was re-staged for every
  // (pair tile, row tile) before: 10% of all stall samples); the small B tile (64 chains x 32 pairs) is
  // rebuilt from Gs in shared memory for every round instead.
  for (int nt = 0; nt < NT; ++nt) {
    const int n0 = nt * 128;
    __syncthreads();  // previous users of xs are done
    for (int e = tid; e < D * 128; e += TC_THREADS) {
      const int i = e >> 7, r = e & 127;
      sm.xs[e] = (n0 + r < N) ? tg.Xt[(size_t)i * tg.ldx + n0 + r] : 0.f;
    }
    __syncthreads();
    for (int kp = 0; kp < KP; ++kp) {
      // B tile: [chain][pair kp*32 ..], off-diagonal pairs count twice in x^T Ginv x
      for (int e = tid; e < TC_CB * TC_KT; e += TC_THREADS) {
        const int cc = e / TC_KT, kk = e - cc * TC_KT;
        const int pr = kp * TC_KT + kk;
        float v = 0.f;
        if (pr < tg.P) v = sm.Gs[(size_t)cc * tg.PS + pr] * (sm.pair_i[pr] == sm.pair_j[pr] ? 1.f : 2.f);
        float hi, lo;
        tc_split(v, hi, lo);
        const int off = tc_off(cc, kk);
        *(float*)(sm.B_hi + off) = hi;
        *(float*)(sm.B_lo + off) = lo;
      }
      {
        const int row = tid & 127, kh = (tid >> 7) * 16;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const int pr = kp * TC_KT + kh + kk;
          float a = 0.f;
          if (pr < tg.P) a = sm.xs[sm.pair_i[pr] * 128 + row] * sm.xs[sm.pair_j[pr] * 128 + row];
          float hi, lo;
          tc_split(a, hi, lo);
          const int off = tc_off(row, kh + kk);
          *(float*)(sm.A_hi + off) = hi;
          *(float*)(sm.A_lo + off) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t td = sm.tmem + (uint32_t)(nt * TC_CB);
#pragma unroll
        for (int k8 = 0; k8 < TC_KT / 8; ++k8) {
          const uint32_t adv = (uint32_t)k8 * 2u * TC_LBO;
          const uint32_t acc0 = (kp > 0 || k8 > 0) ? 1u : 0u;
          tc_mma(td, tc_desc(a_hi_s + adv), tc_desc(b_hi_s + adv), acc0);
          tc_mma(td, tc_desc(a_hi_s + adv), tc_desc(b_lo_s + adv), 1u);
          tc_mma(td, tc_desc(a_lo_s + adv), tc_desc(b_hi_s + adv), 1u);
        }
      }
      tc_commit_and_wait(sm);  // operands are single-buffered
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue per data-row tile: t[n][c] = w'(h - u^2) (w' recomputed from eta), then dT += X^T t
  float* tbuf = (float*)sm.A_hi;  // [128 rows][TC_CB + 1]: the A tiles are free now (32 KB + 16 KB lo)
  const int TS = TC_CB + 1;
  float dacc[7];  // this thread's share of dT: outputs e = tid + 256*k (< TC_CB * D)
#pragma unroll
  for (int k = 0; k < 7; ++k) dacc[k] = 0.f;
  for (int nt = 0; nt < NT; ++nt) {
    const int n0 = nt * 128;
    __syncthreads();
    for (int e = tid; e < D * 128; e += TC_THREADS) {
      const int i = e >> 7, r = e & 127;
      sm.xs[e] = (n0 + r < N) ? tg.Xt[(size_t)i * tg.ldx + n0 + r] : 0.f;
    }
    __syncthreads();
    {
      const int row = (warp & 3) * 32 + lane, cb = (warp >> 2) * 32;
      float x[DP];
#pragma unroll
      for (int i = 0; i < DP; ++i) x[i] = (i < D) ? sm.xs[i * 128 + row] : 0.f;
      const bool live = n0 + row < N;
#pragma unroll 1
      for (int c8 = 0; c8 < 32; c8 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(sm.tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(nt * TC_CB + cb + c8)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int cc = cb + c8 + e;
          float eta = 0.f, u = 0.f;
          const float4* q4 = (const float4*)(qv + cc * TC_DS);
          const float4* w4 = (const float4*)(sm.w + cc * TC_DS);
#pragma unroll
          for (int i4 = 0; i4 < DP / 4; ++i4) {
            const float4 qq = q4[i4], ww = w4[i4];
            eta = fmaf(x[4 * i4], qq.x, eta); eta = fmaf(x[4 * i4 + 1], qq.y, eta);
            eta = fmaf(x[4 * i4 + 2], qq.z, eta); eta = fmaf(x[4 * i4 + 3], qq.w, eta);
            u = fmaf(x[4 * i4], ww.x, u); u = fmaf(x[4 * i4 + 1], ww.y, u);
            u = fmaf(x[4 * i4 + 2], ww.z, u); u = fmaf(x[4 * i4 + 3], ww.w, u);
          }
          const float s = 1.f / (1.f + expf(-eta));
          const float wp = s * (1.f - s) * (1.f - 2.f * s);
          tbuf[row * TS + cc] = live ? wp * (__uint_as_float(r[e]) - u * u) : 0.f;
        }
      }
    }
    __syncthreads();
    // dT[c][i] += sum_rows X[n, i] t[n][c]; output e -> (c = e % TC_CB, i = e / TC_CB)
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int e = tid + TC_THREADS * k;
      if (e < TC_CB * D) {
        const int cc = e & (TC_CB - 1), i = e >> 6;
        float a = dacc[k];
        const float* xi = sm.xs + i * 128;
#pragma unroll 8
        for (int r = 0; r < 128; ++r) a = fmaf(xi[r], tbuf[r * TS + cc], a);
        dacc[k] = a;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int e = tid + TC_THREADS * k;
    if (e < TC_CB * D) sm.dT[(e & (TC_CB - 1)) * TC_DS + (e >> 6)] = 0.5f * dacc[k];
  }
  __syncthreads();
}

// one evaluation of the fixed-point map for the whole tile; candidates land in (qn, pn) = (sm.red, sm.red + half)
template <int DP>
__device__ void tc_map(const LogRegTC& tg, TCSmem& sm, const float* qe, const float* pe, const float* qi,
                       const float* pi, float he, float* qn, float* pn) {
  tc_pass1<DP>(tg, sm, qe, true);
  tc_factor<DP>(tg, sm, false, pe, sm.w);
  tc_pass2<DP>(tg, sm, qe);
  for (int e = threadIdx.x; e < TC_CB * TC_DS; e += TC_THREADS) {
    const int i = e % TC_DS;
    if (i < tg.D) {
      const float a = qi[e], b = pi[e];
      qn[e] = fmaf(he, sm.w[e], a);
      pn[e] = fmaf(-he, sm.dT[e] - sm.g[e], b);
    }
  }
  __syncthreads();
}

// nrm[c] = max_i max(|qa - qb|, |pa - pb|) (inf when not finite)
__device__ void tc_norm(const LogRegTC& tg, TCSmem& sm, const float* qa, const float* pa, const float* qb, const float* pb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < TC_CB; c += TC_THREADS / 32) {
    float mx = 0.f;
    bool nan = false;
    if (lane < tg.D) {
      const float dq = fabsf(qa[c * TC_DS + lane] - qb[c * TC_DS + lane]);
      const float dp = fabsf(pa[c * TC_DS + lane] - pb[c * TC_DS + lane]);
      nan = isnan(dq) || isnan(dp);
      mx = fmaxf(dq, dp);
    }
    if (nan) mx = __int_as_float(0x7f800000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sm.nrm[c] = mx;
  }
  __syncthreads();
}

static size_t tc_smem_bytes(int D, int PS) {
  size_t b = 2 * TC_A_BYTES + 2 * TC_B_BYTES;           // operand tiles
  b += sizeof(float) * (size_t)D * 128;                 // xs (pass 2 size)
  b += sizeof(float) * (size_t)TC_CB * PS;              // Gs
  b += sizeof(float) * 7 * TC_CB * TC_DS;               // q p q0 p0 g w dT(=z)
  b += sizeof(float) * 2 * TC_CB * TC_DS;               // candidate iterate
  b += sizeof(float) * 8 * TC_CB;                       // per-chain scalars
  b += sizeof(short) * 2 * (size_t)PS;
  return b + 1024;
}

template <int DP>
__global__ void __launch_bounds__(TC_THREADS, 1) rmhmc_logreg_tc_kernel(const TransArgs a, const LogRegTC tg) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  TCSmem sm;
  {
    unsigned char* b = tc_raw;
    sm.A_hi = b; b += TC_A_BYTES;
    sm.A_lo = b; b += TC_A_BYTES;
    sm.B_hi = b; b += TC_B_BYTES;
    sm.B_lo = b; b += TC_B_BYTES;
    float* f = (float*)b;
    sm.xs = f; f += (size_t)tg.D * 128;
    sm.Gs = f; f += (size_t)TC_CB * tg.PS;
    sm.q = f; f += TC_CB * TC_DS; sm.p = f; f += TC_CB * TC_DS; sm.q0 = f; f += TC_CB * TC_DS; sm.p0 = f; f += TC_CB * TC_DS;
    sm.g = f; f += TC_CB * TC_DS; sm.w = f; f += TC_CB * TC_DS; sm.dT = f; sm.z = f; f += TC_CB * TC_DS;
    sm.red = f; f += 2 * TC_CB * TC_DS;
    sm.lp = f; f += TC_CB; sm.logdet = f; f += TC_CB; sm.nrm = f; f += TC_CB; sm.H0 = f; f += TC_CB;
    f += 4 * TC_CB;
    sm.pair_i = (short*)f;
    sm.pair_j = sm.pair_i + tg.PS;
  }
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int any_s;
  __shared__ int iters_s[TC_CB], nfp_s[TC_CB], active_s[TC_CB], straggler_s[TC_CB];
  const int tid = threadIdx.x, warp = tid >> 5, D = tg.D;
  sm.mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
  sm.phase = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm.mbar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int pr = tid; pr < tg.PS; pr += TC_THREADS) {
    int i = 0, rem = pr;
    if (pr < tg.P) {
      while (rem >= D - i) { rem -= D - i; ++i; }
      sm.pair_i[pr] = (short)i;
      sm.pair_j[pr] = (short)(i + rem);
    } else {
      sm.pair_i[pr] = 0;
      sm.pair_j[pr] = 0;
    }
  }
  for (int e = tid; e < 9 * TC_CB * TC_DS; e += TC_THREADS) sm.q[e] = 0.f;  // q..dT and the candidates are contiguous: clear padding
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  sm.tmem = tmem_base_s;

  const float tol = (float)a.fp_tol, div_tol = (float)a.fp_div_tol;
  const long long T = a.ks.keys ? 1 : a.ks.num_transitions;
  const long long ntiles = (a.C + TC_CB - 1) / TC_CB;
  float* qn = sm.red;
  float* pn = sm.red + TC_CB * TC_DS;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long c0 = tile * TC_CB;
    for (long long it = 0; it < T; ++it) {
      const long long t = a.ks.first_transition + it;
      const float* spos = (const float*)(it == 0 ? a.in_pos : a.out_pos);
      const float* slogp = (const float*)(it == 0 ? a.in_logp : a.out_logp);
      const float* sgrad = (const float*)(it == 0 ? a.in_grad : a.out_grad);
      // load positions, draw z (one thread per (chain, element))
      for (int e = tid; e < TC_CB * TC_DS; e += TC_THREADS) {
        const int cc = e / TC_DS, i = e - cc * TC_DS;
        const long long chain = (c0 + cc < a.C) ? c0 + cc : a.C - 1;  // padded chains replicate the last one (never written)
        float qv = 0.f, zv = 0.f;
        if (i < D) {
          qv = spos[chain * D + i];
          if (a.opts.noise_override != nullptr) zv = ((const float*)a.opts.noise_override)[chain * D + i];
          else {
            U2 key = transition_key(a, chain, t);
            U2 k_m, k_a;
            split2(a.mode, key, k_m, k_a);
            zv = bits_to_normal(random_bits_elem(a.mode, k_m, (uint32_t)i, (uint32_t)D));
          }
          if (a.info.noise && c0 + cc < a.C) ((float*)a.info.noise)[chain * D + i] = zv;
        }
        sm.q[e] = qv;
        sm.z[e] = zv;
      }
      __syncthreads();
      // start of the transition: metric, momentum p = L z, w = G^-1 p, H0
      tc_pass1<DP>(tg, sm, sm.q, true);
      tc_factor<DP>(tg, sm, true, sm.p, sm.w);
      if (tid < TC_CB) {
        const long long chain = (c0 + tid < a.C) ? c0 + tid : a.C - 1;
        float pw = 0.f;
        for (int i = 0; i < D; ++i) pw = fmaf(sm.p[tid * TC_DS + i], sm.w[tid * TC_DS + i], pw);
        sm.H0[tid] = -slogp[chain] + 0.5f * pw + 0.5f * sm.logdet[tid] + 0.91893853320467274178f * (float)D;
        iters_s[tid] = 0;
        straggler_s[tid] = 0;
      }
      if (a.info.momentum) {
        for (int e = tid; e < TC_CB * TC_DS; e += TC_THREADS) {
          const int cc = e / TC_DS, i = e - cc * TC_DS;
          if (i < D && c0 + cc < a.C) ((float*)a.info.momentum)[(c0 + cc) * D + i] = sm.p[e];
        }
      }
      __syncthreads();
      float eps = (float)a.step_size;  // per-chain step sizes are not supported by the lock-step tile kernel
      const float he = 0.5f * eps;
      for (int s = 0; s < a.num_steps; ++s) {
        for (int e = tid; e < TC_CB * TC_DS; e += TC_THREADS) { sm.q0[e] = sm.q[e]; sm.p0[e] = sm.p[e]; }
        __syncthreads();
        tc_map<DP>(tg, sm, sm.q0, sm.p0, sm.q0, sm.p0, he, sm.q, sm.p);
        tc_norm(tg, sm, sm.q, sm.p, sm.q0, sm.p0);
        if (tid < TC_CB) nfp_s[tid] = 0;
        __syncthreads();
        for (;;) {
          if (tid == 0) any_s = 0;
          __syncthreads();
          if (tid < TC_CB) {
            const float nr = sm.nrm[tid];
            int go = (nfp_s[tid] < a.fp_max_iters) && (nr < __int_as_float(0x7f800000)) && (nr < div_tol) && (nr > tol);
            if (go && nfp_s[tid] >= a.lock_cap) {  // heavy-tailed chain: freeze it here, re-run it chain by chain
              straggler_s[tid] = 1;
              go = 0;
            }
            active_s[tid] = go;
            if (go) any_s = 1;
          }
          __syncthreads();
          if (!any_s) break;
          tc_map<DP>(tg, sm, sm.q, sm.p, sm.q0, sm.p0, he, qn, pn);
          // masked commit (vmapped while_loop): only still-active chains take the new iterate and norm
          {
            const int lane = tid & 31;
            for (int c = warp; c < TC_CB; c += TC_THREADS / 32) {
              float mx = 0.f;
              bool nan = false;
              if (lane < D) {
                const float dq = fabsf(qn[c * TC_DS + lane] - sm.q[c * TC_DS + lane]);
                const float dp = fabsf(pn[c * TC_DS + lane] - sm.p[c * TC_DS + lane]);
                nan = isnan(dq) || isnan(dp);
                mx = fmaxf(dq, dp);
              }
              if (nan) mx = __int_as_float(0x7f800000);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
              if (active_s[c]) {
                if (lane < D) { sm.q[c * TC_DS + lane] = qn[c * TC_DS + lane]; sm.p[c * TC_DS + lane] = pn[c * TC_DS + lane]; }
                if (lane == 0) { sm.nrm[c] = mx; nfp_s[c] += 1; }
              }
            }
          }
          __syncthreads();
        }
        if (tid < TC_CB) iters_s[tid] += nfp_s[tid];
        tc_map<DP>(tg, sm, sm.q, sm.p, sm.q, sm.p, he, qn, pn);  // explicit update from the midpoint
        for (int e = tid; e < TC_CB * TC_DS; e += TC_THREADS) { sm.q[e] = qn[e]; sm.p[e] = pn[e]; }
        __syncthreads();
      }
      // end state: log-density, gradient, velocity, energy, accept
      tc_pass1<DP>(tg, sm, sm.q, true);
      tc_factor<DP>(tg, sm, false, sm.p, sm.w);
      if (tid < TC_CB && c0 + tid < a.C && straggler_s[tid]) {
        a.work_list[atomicAdd(a.work_count, 1)] = (int)(c0 + tid);
      }
      if (tid < TC_CB && c0 + tid < a.C && !straggler_s[tid]) {
        const long long chain = c0 + tid;
        float pw = 0.f;
        for (int i = 0; i < D; ++i) pw = fmaf(sm.p[tid * TC_DS + i], sm.w[tid * TC_DS + i], pw);
        const float lp = sm.lp[tid];
        const float H1 = -lp + 0.5f * pw + 0.5f * sm.logdet[tid] + 0.91893853320467274178f * (float)D;
        U2 key = transition_key(a, chain, t);
        U2 k_m, k_a;
        split2(a.mode, key, k_m, k_a);
        MH<float> mh = metropolis<float>(a, k_a, chain, sm.H0[tid], H1);
        active_s[tid] = mh.accept;
        store_scalar<float>(a.info.acceptance_rate, chain, mh.p_accept);
        if (a.info.is_accepted) a.info.is_accepted[chain] = mh.accept;
        if (a.info.is_divergent) a.info.is_divergent[chain] = mh.divergent;
        store_scalar<float>(a.info.energy, chain, H1);
        store_scalar<float>(a.info.proposal_logdensity, chain, lp);
        store_scalar<float>(a.info.proposal_weight, chain, mh.weight);
        store_scalar<float>(a.info.initial_energy, chain, sm.H0[tid]);
        store_scalar<float>(a.info.accept_uniform, chain, mh.u);
        if (a.info.fp_iters) a.info.fp_iters[chain] = iters_s[tid];
        if (a.opts.sample_accept != nullptr) ((float*)a.opts.sample_accept)[it * a.C + chain] = mh.p_accept;
        ((float*)a.out_logp)[chain] = mh.accept ? lp : slogp[chain];
      }
      __syncthreads();
      for (int e = tid; e < TC_CB * TC_DS; e += TC_THREADS) {
        const int cc = e / TC_DS, i = e - cc * TC_DS;
        if (i < D && c0 + cc < a.C && !straggler_s[cc]) {
          const long long o = (c0 + cc) * D + i;
          if (a.info.proposal_position) ((float*)a.info.proposal_position)[o] = sm.q[e];
          if (a.info.proposal_momentum) ((float*)a.info.proposal_momentum)[o] = -sm.p[e];
          if (a.info.proposal_velocity) ((float*)a.info.proposal_velocity)[o] = -sm.w[e];
          if (a.info.proposal_logdensity_grad) ((float*)a.info.proposal_logdensity_grad)[o] = sm.g[e];
          const float qo = active_s[cc] ? sm.q[e] : spos[o];
          const float go = active_s[cc] ? sm.g[e] : sgrad[o];
          ((float*)a.out_pos)[o] = qo;
          ((float*)a.out_grad)[o] = go;
          if (a.opts.samples != nullptr) ((float*)a.opts.samples)[(it * a.C) * D + o] = qo;
        }
      }
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem), "n"(512));
}

// returns GB200_ERR_UNSUPPORTED when the shape does not fit this kernel (caller falls back)
int launch_rmhmc_logreg_tc(const TransArgs& a, const gb200_target_desc& t, cudaStream_t s) {
  LogRegTC tg;
  tg.Xt = (const float*)t.vec0;
  tg.y = (const float*)t.y;
  tg.N = (int)t.N;
  tg.D = t.D;
  tg.ldx = (int)t.params[1];
  tg.alpha = (float)t.params[0];
  tg.P = tg.D * (tg.D + 1) / 2;
  tg.PS = (tg.P + TC_KT - 1) / TC_KT * TC_KT;
  const int MT = (tg.P + 127) / 128, NT = (tg.N + 127) / 128;
  if (tg.D > 28 || tg.D < 2 || MT * TC_CB > 512 || NT * TC_CB > 512) return GB200_ERR_UNSUPPORTED;
  if (a.step_size_per_chain != nullptr || a.opts.dual_averaging != nullptr) return GB200_ERR_UNSUPPORTED;
  if (a.work_list == nullptr || a.work_count == nullptr) return GB200_ERR_UNSUPPORTED;
  const size_t smem = tc_smem_bytes(tg.D, tg.PS);
  if (smem > 227 * 1024) return GB200_ERR_UNSUPPORTED;
  const long long ntiles = (a.C + TC_CB - 1) / TC_CB;
  const int grid = (int)(ntiles < 148 ? ntiles : 148);
  const int dp = (tg.D + 3) / 4 * 4;
#define GB_TC(DPV)                                                                                                      \
  if (dp == DPV) {                                                                                                      \
    cudaError_t e = cudaFuncSetAttribute(rmhmc_logreg_tc_kernel<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) { set_error("logreg_tc: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }                 \
    rmhmc_logreg_tc_kernel<DPV><<<grid, TC_THREADS, smem, s>>>(a, tg);                                                  \
    GB_CHECK_LAUNCH();                                                                                                  \
    return GB200_OK;                                                                                                    \
  }
  GB_TC(4) GB_TC(8) GB_TC(20) GB_TC(28)  // other D run on the CTA-per-chain kernel (compile time)
#undef GB_TC
  return GB200_ERR_UNSUPPORTED;
}

}  // namespace gb
