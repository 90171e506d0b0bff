// rmhmc on Bayesian logistic regression (Fisher metric): the lock-step "rolling batch" sampler.
//
// Reference semantics (under jax.vmap): rmhmc/rmhmc.py:131-174 (kernel), :416-462 (proposal, flip),
// rmhmc/integrators.py:53-89 (solve_fixed_point_iteration), :92-156 (implicit midpoint),
// rmhmc/metrics.py:42-129 (momentum draw, kinetic energy, G^-1 p), mcmc/proposal.py:87-121,168-185 (accept),
// optimizers/dual_averaging.py:101-123 (fused step-size adaptation), examples/funnel/main.py:7-25 (key tree).
//
// Every evaluation of the implicit-midpoint map costs two D^2 N products per chain (metric G = X^T diag(w) X and
// the quadratic forms x_n^T (G^-1 - v v^T) x_n); both are shared-operand GEMMs over the chain dimension and run on
// the warp-specialised tcgen05 kernels of fisher_tc.cu.  A vmapped while_loop would make every chain wait for the
// slowest chain of every integrator step (the fixed-point count is heavy-tailed).  Chains are independent, so
// instead each chain carries its own small state machine
//     FIRST0 (draw p = chol(G) z, H0, x1 = f(x0)) -> ITER* -> EXPL -> [FIRST -> ITER* -> EXPL] x (L-1) -> END
// and ONE round evaluates the map once for every chain that is not finished, whatever its phase:
//     ls_weights  : eta = X q, s, w = s(1-s)  -> W operand tiles (TF32 hi / lo), s, log-density partials
//     metric GEMM : vec(G) = Z^T W                                            (tcgen05, 3xTF32)
//     ls_factor   : blocked Cholesky in registers, log det, [p = L z], G^-1, v = G^-1 p, A' = G^-1 - v v^T
//     quad_b      : A' -> B operand tiles
//     quad GEMM   : x_n^T A' x_n, fused epilogue R = 1/2 w' quad - (y - s), partial X^T R  (tcgen05, 3xTF32)
//     ls_reduce   : dH/dq = X^T R (+ alpha q), log-density
//     ls_advance  : candidate iterate, convergence test, phase transition, MH accept, outputs, dual averaging,
//                   next transition's key / noise (fused multi-transition launches keep the batch full)
//     ls_compact  : list of the chains still running, loop condition
// The round is the body of a CUDA-graph WHILE node (cudaGraphCondTypeWhile): the whole multi-transition launch is
// asynchronous on the caller's stream with no host round trip.  A chain's numbers never depend on which other
// chains ride in its tile, so a fused launch equals the same transitions launched one by one, bit for bit.
#include <stdio.h>
#include <new>
#include "fisher_tc.cuh"

namespace gb {

struct LsDims {
  int N, D, ldx, ldn, P, PS, ktF, ktQ, mtQ, BS, NB, nblk, fthreads, fwarp, fsmem_floats, ksplit;
  long long Ccap, ctiles;
  float alpha;
};

struct LsCall {  // per-launch arguments, copied to device memory by ls_start_kernel
  TransArgs a;
  long long T;
  long long max_rounds;
};

struct LsBuf {
  // control
  int* n_active;        // [0] chains still running, [1] rounds executed, [2] chain-evaluations executed (mod 2^31)
  LsCall* call;
  // per chain [Ccap]
  float *q, *p, *qi, *pi, *z, *p0;  // [Ccap, D]
  float *he, *H0;
  unsigned char* phase;
  int *step, *nit, *tcur, *iters;
  // per slot (compacted list of running chains)
  int* idx;
  unsigned char* slot_phase;
  float *v, *logdet, *sbuf, *Gp, *Ap, *parts, *lp_parts, *dHt, *lpt, *qT;
  // operands
  unsigned char *Wt, *Bt;
  float* Xtile;
  short2* pairs;
  const float *Xt, *y;
};

__device__ __forceinline__ float ls_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ls_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Start transition t_rel of chain c (one warp): key, noise, position, step size.
// rmhmc/rmhmc.py:158 (split), util.py:81-82 (normal), examples/funnel/main.py:18,22 (key tree)
__device__ __forceinline__ void ls_begin_transition(const TransArgs& a, const LsBuf& b, int D, int c, long long t_rel, int lane) {
  const long long t = a.ks.first_transition + t_rel;
  const U2 key = transition_key(a, c, t);
  U2 k_m, k_a;
  split2(a.mode, key, k_m, k_a);
  for (int i = lane; i < D; i += 32) {
    float zz;
    if (a.opts.noise_override != nullptr) zz = ((const float*)a.opts.noise_override)[(size_t)c * D + i];
    else zz = bits_to_normal(random_bits_elem(a.mode, k_m, (uint32_t)i, (uint32_t)D));
    b.z[(size_t)c * D + i] = zz;
    if (a.info.noise) ((float*)a.info.noise)[(size_t)c * D + i] = zz;
    const float pos = ((const float*)a.out_pos)[(size_t)c * D + i];
    b.q[(size_t)c * D + i] = pos;
    b.qi[(size_t)c * D + i] = pos;  // x0 of the first step; its momentum half is set by ls_factor_kernel (p = L z)
  }
  if (lane == 0) {
    float eps = (float)a.step_size;
    if (a.opts.dual_averaging != nullptr) eps = expf(((const float*)a.opts.dual_averaging)[(size_t)c * 5]);
    else if (a.step_size_per_chain != nullptr) eps = ((const float*)a.step_size_per_chain)[c];
    b.he[c] = 0.5f * eps;
    b.phase[c] = LS_PH_FIRST0;
    b.step[c] = 0;
    b.nit[c] = 0;
    b.iters[c] = 0;
  }
}

// one warp per chain: the current state becomes the out state (the carry of a fused launch), first transition begins
__global__ void __launch_bounds__(256) ls_start_kernel(const LsCall call, const LsBuf b, const LsDims d) {
  if (blockIdx.x == 0) {
    const int* src = (const int*)&call;
    int* dst = (int*)b.call;
    for (int k = threadIdx.x; k < (int)(sizeof(LsCall) / 4); k += blockDim.x) dst[k] = src[k];
    if (threadIdx.x == 0) { b.n_active[1] = 0; b.n_active[2] = 0; }
  }
  const int lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= d.Ccap) return;
  const TransArgs& a = call.a;
  if (c >= a.C) {
    if (lane == 0) b.phase[c] = LS_PH_DONE;
    return;
  }
  const int D = d.D;
  if (a.in_pos != a.out_pos)
    for (int i = lane; i < D; i += 32) ((float*)a.out_pos)[c * D + i] = ((const float*)a.in_pos)[c * D + i];
  if (a.in_grad != a.out_grad)
    for (int i = lane; i < D; i += 32) ((float*)a.out_grad)[c * D + i] = ((const float*)a.in_grad)[c * D + i];
  if (lane == 0) {
    if (a.in_logp != a.out_logp) ((float*)a.out_logp)[c] = ((const float*)a.in_logp)[c];
    b.tcur[c] = 0;
  }
  __syncwarp();
  ls_begin_transition(a, b, D, (int)c, 0, lane);
}

// stable compaction of the running chains (single CTA) + the WHILE condition of the graph
__global__ void __launch_bounds__(1024) ls_compact_kernel(const LsBuf b, const LsDims d, cudaGraphConditionalHandle handle,
                                                          int use_handle, int count_round) {
  __shared__ int wsum[32];
  __shared__ int total_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Each WARP owns a contiguous run of chains (a multiple of 32 long) and walks it 32 chains at a time, lane = chain:
  // one ballot gives the strip's live count and every lane's rank, the loads and the stores are coalesced, the order
  // stays the chain order.  (One contiguous run per THREAD wrote idx / slot_phase with a 64-byte stride between
  // lanes: 18 us at 16,384 chains on the one SM this kernel runs on.)
  const long long per = ((d.Ccap + 31) / 32 + 31) / 32 * 32;
  const long long lo = (long long)warp * per, hi = lo + per < d.Ccap ? lo + per : d.Ccap;
  constexpr int U = 8;  // strips in flight
  int cnt = 0;
  for (long long base = lo; base < hi; base += 32 * U) {
    unsigned char ph[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long c = base + 32 * u + lane;
      ph[u] = c < hi ? b.phase[c] : (unsigned char)LS_PH_DONE;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) cnt += __popc(__ballot_sync(0xffffffffu, ph[u] != LS_PH_DONE));
  }
  if (lane == 0) wsum[warp] = cnt;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    wsum[lane] = wi - w;
    if (lane == 31) total_s = wi;
  }
  __syncthreads();
  int pos = wsum[warp];
  const unsigned int lt = (1u << lane) - 1u;
  for (long long base = lo; base < hi; base += 32 * U) {
    unsigned char ph[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long c = base + 32 * u + lane;
      ph[u] = c < hi ? b.phase[c] : (unsigned char)LS_PH_DONE;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool live = ph[u] != LS_PH_DONE;
      const unsigned int m = __ballot_sync(0xffffffffu, live);
      if (live) {
        const int o = pos + __popc(m & lt);
        b.idx[o] = (int)(base + 32 * u + lane);
        b.slot_phase[o] = ph[u];
      }
      pos += __popc(m);
    }
  }
  if (tid == 0) {
    const int total = total_s;
    b.n_active[0] = total;
    int rounds = b.n_active[1];
    if (count_round) {
      rounds += 1;
      b.n_active[1] = rounds;
    }
    if (use_handle) cudaGraphSetConditional(handle, (total > 0 && (long long)rounds < b.call->max_rounds) ? 1u : 0u);
  }
}

// positions of the running chains, transposed: qT[i][slot] = q[idx[slot]][i], so that the weights kernel reads them
// coalesced over slots (32 x 32 tiles through shared memory)
__global__ void __launch_bounds__(256) ls_gather_kernel(const LsBuf b, const LsDims d) {
  __shared__ float tile[32][33];
  const long long nact = b.n_active[0];
  const long long j0 = (long long)blockIdx.x * 32;
  if (j0 >= nact) return;
  const int i0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long j = j0 + r;
    const int i = i0 + tx;
    tile[r][tx] = (j < nact && i < d.D) ? b.q[(size_t)b.idx[j] * d.D + i] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    const long long j = j0 + tx;
    if (i < d.D && j < d.Ccap) b.qT[(size_t)i * d.Ccap + j] = tile[tx][r];
  }
}

// eta = X q, s = sigmoid(eta), w = s(1-s) for the running chains: W operand tiles of the metric GEMM (TF32 hi / lo in
// the UMMA canonical layout, one 32 KB block per (256-chain tile, 16-row K tile), see fisher_weights_kernel), s[slot, n],
// and for chains at the end of their trajectory the log-density partial sums over the K tile's 16 data rows.
// One block = one (chain tile, K tile); thread = one chain, all 16 data rows of the tile: the X tile [D][16] is staged
// in shared memory and read as warp-uniform float4 broadcasts, the positions come coalesced from qT -- 5 loads per 16
// FMAs with no uncoalesced access (the first version, thread = (chain, 4 rows), issued 8 L2-bound loads per 16 FMAs).
// Each thread carries LS_WCPT chains (slots row and row + 128 of the tile): the broadcast X reads, one shared-memory
// wavefront per 16 FMAs of ONE chain, were what bound this kernel (FMA pipe 39 %); two chains per thread halve them.
constexpr int LS_WCPT = 2;
constexpr int LS_WTHREADS = FT_N / LS_WCPT;
__global__ void __launch_bounds__(LS_WTHREADS) ls_weights_kernel(const LsBuf b, const LsDims d) {
  extern __shared__ __align__(16) float wsm[];  // [D][FT_KT]
  const long long nact = b.n_active[0];
  const int ktiles = d.ktF;
  const long long tile = blockIdx.x;
  const long long ct = tile / ktiles;
  if (ct * FT_N >= nact) return;
  const int kt = (int)(tile - ct * ktiles);
  const int N = d.N, D = d.D, n0 = kt * FT_KT;
  for (int e = threadIdx.x; e < D * (FT_KT / 4); e += LS_WTHREADS) {
    const int i = e / (FT_KT / 4), k4 = e - i * (FT_KT / 4);
    // ldx % 4 == 0 and columns in [N, ldx) are zero; K tiles may reach beyond ldx
    const int n = n0 + 4 * k4;
    *(float4*)(wsm + i * FT_KT + 4 * k4) = n < d.ldx ? __ldg((const float4*)(b.Xt + (size_t)i * d.ldx + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const long long j0 = ct * FT_N + threadIdx.x;
  bool live[LS_WCPT], end[LS_WCPT];
#pragma unroll
  for (int c = 0; c < LS_WCPT; ++c) {
    const long long j = j0 + c * LS_WTHREADS;
    live[c] = j < nact;
    end[c] = live[c] && b.slot_phase[j] == LS_PH_END;
  }
  float eta[LS_WCPT][FT_KT];
#pragma unroll
  for (int c = 0; c < LS_WCPT; ++c)
#pragma unroll
    for (int k = 0; k < FT_KT; ++k) eta[c][k] = 0.f;
  if (live[0]) {  // slots are compacted: live[c] implies live[0]
    // positions are fetched CH features ahead (CH independent loads in flight per chain); addresses advance by
    // pointer bumps and constant offsets (the first version recomputed them per feature: 2.6 instructions per FMA)
    constexpr int CH = 10;
    const float* tp = b.qT + j0;
    const float* xp = wsm;
    const size_t cs = (size_t)d.Ccap;
    int i = 0;
    for (; i + CH <= D; i += CH) {
      float t[LS_WCPT][CH];
#pragma unroll
      for (int u = 0; u < CH; ++u)
#pragma unroll
        for (int c = 0; c < LS_WCPT; ++c) t[c][u] = live[c] ? __ldg(tp + u * cs + c * LS_WTHREADS) : 0.f;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
#pragma unroll
        for (int k4 = 0; k4 < FT_KT / 4; ++k4) {
          const float4 x = *(const float4*)(xp + u * FT_KT + 4 * k4);
#pragma unroll
          for (int c = 0; c < LS_WCPT; ++c) {
            eta[c][4 * k4 + 0] = fmaf(x.x, t[c][u], eta[c][4 * k4 + 0]);
            eta[c][4 * k4 + 1] = fmaf(x.y, t[c][u], eta[c][4 * k4 + 1]);
            eta[c][4 * k4 + 2] = fmaf(x.z, t[c][u], eta[c][4 * k4 + 2]);
            eta[c][4 * k4 + 3] = fmaf(x.w, t[c][u], eta[c][4 * k4 + 3]);
          }
        }
      }
      tp += CH * cs;
      xp += CH * FT_KT;
    }
    for (; i < D; ++i) {
      float t[LS_WCPT];
#pragma unroll
      for (int c = 0; c < LS_WCPT; ++c) t[c] = live[c] ? __ldg(tp + c * LS_WTHREADS) : 0.f;
#pragma unroll
      for (int k4 = 0; k4 < FT_KT / 4; ++k4) {
        const float4 x = *(const float4*)(xp + 4 * k4);
#pragma unroll
        for (int c = 0; c < LS_WCPT; ++c) {
          eta[c][4 * k4 + 0] = fmaf(x.x, t[c], eta[c][4 * k4 + 0]);
          eta[c][4 * k4 + 1] = fmaf(x.y, t[c], eta[c][4 * k4 + 1]);
          eta[c][4 * k4 + 2] = fmaf(x.z, t[c], eta[c][4 * k4 + 2]);
          eta[c][4 * k4 + 3] = fmaf(x.w, t[c], eta[c][4 * k4 + 3]);
        }
      }
      tp += cs;
      xp += FT_KT;
    }
  }
#pragma unroll
  for (int c = 0; c < LS_WCPT; ++c) {
    const int row = threadIdx.x + c * LS_WTHREADS;
    const long long j = j0 + c * LS_WTHREADS;
    float lp = 0.f;
    unsigned char* base = b.Wt + (size_t)tile * (2 * FT_B_BYTES) + (size_t)(row >> 3) * FT_SBO + (row & 7) * 16;
#pragma unroll
    for (int k4 = 0; k4 < FT_KT / 4; ++k4) {
      float sg[4], w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + 4 * k4 + e;
        const float et = eta[c][4 * k4 + e];
        sg[e] = 1.f / (1.f + expf(-et));
        w[e] = (live[c] && n < N) ? sg[e] * (1.f - sg[e]) : 0.f;
        if (end[c] && n < N) lp += __ldg(b.y + n) * et - (fmaxf(et, 0.f) + log1pf(expf(-fabsf(et))));  // jnp.logaddexp(0, eta)
      }
      if (live[c]) {  // sT[data row][chain slot]: lanes = consecutive slots
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n0 + 4 * k4 + e < N) b.sbuf[(size_t)(n0 + 4 * k4 + e) * d.ldn + j] = sg[e];
      }
      float4 hi, lo;
      ft_split(w[0], hi.x, lo.x); ft_split(w[1], hi.y, lo.y); ft_split(w[2], hi.z, lo.z); ft_split(w[3], hi.w, lo.w);
      *(float4*)(base + k4 * FT_LBO) = hi;
      *(float4*)(base + k4 * FT_LBO + FT_B_BYTES) = lo;
    }
    if (end[c]) b.lp_parts[(size_t)kt * d.Ccap + j] = lp;
  }
}

// ---- per-chain dense algebra: blocked Cholesky / inverse with the matrix in REGISTERS ---------------------------
// CTA = one chain.  The lower triangle of the D x D metric is cut into BS x BS blocks; thread t owns block
// (bi >= bj) in registers for the whole factorisation.  Right-looking blocked Cholesky: per block column the
// diagonal owner factors its block (and inverts it), the panel owners do their triangular solve against that
// inverse, everybody else applies one BS x BS x BS rank update from two panel blocks read from shared memory
// (k-major, float4): two CTA barriers per BLOCK column instead of three per column, 4 FMA per shared-memory word
// instead of 1/3.  L^-1 (block forward substitution, one barrier per block diagonal) and G^-1 = L^-T L^-1
// (independent block products) follow the same pattern.  rmhmc/metrics.py:45-58 (p = chol(G) z), :60-74 (log det),
// :120-127 (v = G^-1 p).
__device__ __forceinline__ int ls_blk(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }

// WARP: the block triangle fits one warp (nblk <= 32, i.e. D <= 28 with BS = 4): one WARP per chain, four chains per
// CTA, __syncwarp instead of CTA barriers (c4: D = 25 used 28 of a 64-thread CTA's lanes with 20 CTA barriers).
// MAXT / MINB: D <= 120 with BS = 8 needs at most 128 threads; capping the registers at 128 (four CTAs per SM, what
// shared memory allows at D = 100) instead of the 129 the compiler picks raises the residency from 3 to 4 chains per SM.
template <int BS, bool WARP, int MAXT = (BS == 8 ? 160 : 128), int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB) ls_factor_kernel(const LsBuf b, const LsDims d) {
  extern __shared__ __align__(16) float fsm_all[];
  const long long j = WARP ? (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : (long long)blockIdx.x;
  if (j >= b.n_active[0]) return;  // whole warp / whole CTA
  const int c = b.idx[j];
  const int ph = b.slot_phase[j];
  const int D = d.D, NB = d.NB, nblk = d.nblk;
  const int tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x, nthr = WARP ? 32 : (int)blockDim.x;
  constexpr int ST = BS * BS + 4;
  float* fsm = fsm_all + (WARP ? (size_t)(threadIdx.x >> 5) * d.fsmem_floats : 0);
  auto sync = [&]() { if (WARP) __syncwarp(); else __syncthreads(); };
  float* Ls = fsm;               // [nblk][ST] L blocks, k-major: Ls[blk][k BS + a] = L[BS bi + a][BS bj + k]
  float* Li = Ls + nblk * ST;    // [nblk][ST] L^-1 blocks, row-major
  float* Dv = Li + nblk * ST;    // [NB][ST]   inverses of the diagonal blocks of L, row-major
  const int DP = NB * BS;
  float* ps = Dv + NB * ST;      // [DP] p (or z)
  float* vs = ps + DP;           // [DP]
  float* logd = vs + DP;         // [NB]
  int bi = 0, bj = 0;
  const bool has = tid < nblk;
  if (has) {
    int rem = tid;
    while (rem >= bi + 1) { rem -= bi + 1; ++bi; }
    bj = rem;
  }
  float A[BS][BS];
  if (d.ksplit > 1) {
    // split-K parts of the metric GEMM (at most 4): summed in a fixed order by a coalesced pass, in place in part 0
    // (added inside the block loads below, the strided 32-byte reads of three more parts cost 170 us at c5's shape)
    float* G0 = b.Gp + (size_t)j * d.PS;
    const size_t zs = (size_t)d.Ccap * d.PS;
    for (int m4 = 4 * tid; m4 < d.PS; m4 += 4 * nthr) {
      float4 v = *(const float4*)(G0 + m4);
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 v1 = *(const float4*)(G0 + zs + m4);
      const float4 v2 = d.ksplit > 2 ? *(const float4*)(G0 + 2 * zs + m4) : z4;
      const float4 v3 = d.ksplit > 3 ? *(const float4*)(G0 + 3 * zs + m4) : z4;
      v.x = ((v.x + v1.x) + v2.x) + v3.x;
      v.y = ((v.y + v1.y) + v2.y) + v3.y;
      v.z = ((v.z + v1.z) + v2.z) + v3.z;
      v.w = ((v.w + v1.w) + v2.w) + v3.w;
      *(float4*)(G0 + m4) = v;
    }
    sync();
  }
  {
    const float* Gp = b.Gp + (size_t)j * d.PS;
#pragma unroll
    for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < BS; ++b_) {
        const int i = BS * bi + a_, jj = BS * bj + b_;
        float v = 0.f;
        if (has) {
          if (i < D && jj <= i) {
            const int m = jj * D - jj * (jj - 1) / 2 + (i - jj);
            v = Gp[m];
          }
          else if (i == jj) v = 1.f;  // identity padding
        }
        A[a_][b_] = v;
      }
  }
  for (int i = tid; i < DP; i += nthr)
    ps[i] = i < D ? (ph == LS_PH_FIRST0 ? b.z[(size_t)c * D + i] : b.p[(size_t)c * D + i]) : 0.f;

  // ---- Cholesky
  for (int kb = 0; kb < NB; ++kb) {
    if (has && bi == kb && bj == kb) {
      // this thread works alone while the rest of the chain's threads wait at the barrier: one reciprocal per
      // pivot (no divisions below), and ONE logarithm per block (of the product of its pivots)
      float rd[BS];
#pragma unroll
      for (int k = 0; k < BS; ++k) {
        const float dk = sqrtf(A[k][k]);
        A[k][k] = dk;
        const float rk = 1.f / dk;
        rd[k] = rk;
#pragma unroll
        for (int a_ = k + 1; a_ < BS; ++a_) A[a_][k] *= rk;
#pragma unroll
        for (int a_ = k + 1; a_ < BS; ++a_)
#pragma unroll
          for (int b_ = k + 1; b_ <= a_; ++b_) A[a_][b_] = fmaf(-A[a_][k], A[b_][k], A[a_][b_]);
      }
      float W[BS][BS];  // inverse of the lower-triangular block
#pragma unroll
      for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
        for (int b_ = 0; b_ < BS; ++b_) W[a_][b_] = 0.f;
#pragma unroll
      for (int cc = 0; cc < BS; ++cc) {
        W[cc][cc] = rd[cc];
#pragma unroll
        for (int r = cc + 1; r < BS; ++r) {
          float s = 0.f;
#pragma unroll
          for (int k = cc; k < r; ++k) s = fmaf(A[r][k], W[k][cc], s);
          W[r][cc] = -s * rd[r];
        }
      }
      float* lo = Ls + ls_blk(kb, kb) * ST;
      float* dv = Dv + kb * ST;
      float pr0 = 1.f, pr1 = 1.f;  // pivots of the first / second half of the block (two products: no overflow for |pivot| < 1e9)
#pragma unroll
      for (int a_ = 0; a_ < BS; ++a_) {
#pragma unroll
        for (int k = 0; k < BS; ++k) {
          lo[k * BS + a_] = (k <= a_) ? A[a_][k] : 0.f;
          dv[a_ * BS + k] = W[a_][k];
        }
        if (BS * kb + a_ < D) {
          if (a_ < BS / 2) pr0 *= A[a_][a_]; else pr1 *= A[a_][a_];
        }
      }
      logd[kb] = logf(pr0) + logf(pr1);
    }
    sync();
    if (has && bj == kb && bi > kb) {  // panel: A <- A L_kk^-T, in place (descending column index)
      const float* W = Dv + kb * ST;
#pragma unroll
      for (int b_ = BS - 1; b_ >= 0; --b_) {
        float wr[BS];
#pragma unroll
        for (int k = 0; k < BS; ++k) wr[k] = W[b_ * BS + k];
#pragma unroll
        for (int a_ = 0; a_ < BS; ++a_) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k <= b_; ++k) s = fmaf(A[a_][k], wr[k], s);
          A[a_][b_] = s;
        }
      }
      float* lo = Ls + ls_blk(bi, kb) * ST;
#pragma unroll
      for (int k = 0; k < BS; ++k)
#pragma unroll
        for (int a_ = 0; a_ < BS; ++a_) lo[k * BS + a_] = A[a_][k];
    }
    sync();
    if (has && bj > kb) {  // trailing update: A -= L[bi][kb] L[bj][kb]^T
      const float* La = Ls + ls_blk(bi, kb) * ST;
      const float* Lb = Ls + ls_blk(bj, kb) * ST;
#pragma unroll
      for (int k = 0; k < BS; ++k) {
        float la[BS], lb[BS];
#pragma unroll
        for (int e = 0; e < BS; e += 4) {
          *(float4*)(la + e) = *(const float4*)(La + k * BS + e);
          *(float4*)(lb + e) = *(const float4*)(Lb + k * BS + e);
        }
#pragma unroll
        for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
          for (int b_ = 0; b_ < BS; ++b_) A[a_][b_] = fmaf(-la[a_], lb[b_], A[a_][b_]);
      }
    }
  }
  sync();

  // ---- momentum draw p = L z (rmhmc/metrics.py:45-58) at the first evaluation of a transition
  if (ph == LS_PH_FIRST0) {
    for (int i = tid; i < D; i += nthr) {
      const int bi_ = i / BS, a_ = i - bi_ * BS;
      float s = 0.f;
      for (int jj = 0; jj <= i; ++jj) {
        const int bj_ = jj / BS, k = jj - bj_ * BS;
        s = fmaf(Ls[ls_blk(bi_, bj_) * ST + k * BS + a_], ps[jj], s);
      }
      vs[i] = s;
    }
    sync();
    for (int i = tid; i < D; i += nthr) {
      const float pv = vs[i];
      ps[i] = pv;
      b.p[(size_t)c * D + i] = pv;
      b.pi[(size_t)c * D + i] = pv;
      b.p0[(size_t)c * D + i] = pv;
    }
    sync();
  }

  // ---- L^-1 by block forward substitution (A is reused as the accumulator S)
#pragma unroll
  for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
    for (int b_ = 0; b_ < BS; ++b_) A[a_][b_] = 0.f;
  if (has && bi == bj) {
    const float* dv = Dv + bi * ST;
    float* li = Li + ls_blk(bi, bi) * ST;
#pragma unroll
    for (int e = 0; e < BS * BS; e += 4) *(float4*)(li + e) = *(const float4*)(dv + e);
  }
  sync();
  for (int m = 0; m < NB - 1; ++m) {
    if (has && bi - bj > m) {
      const int kb = bj + m;
      const float* La = Ls + ls_blk(bi, kb) * ST;   // k-major
      const float* Lk = Li + ls_blk(kb, bj) * ST;   // row-major
#pragma unroll
      for (int k = 0; k < BS; ++k) {
        float la[BS], lk[BS];
#pragma unroll
        for (int e = 0; e < BS; e += 4) {
          *(float4*)(la + e) = *(const float4*)(La + k * BS + e);
          *(float4*)(lk + e) = *(const float4*)(Lk + k * BS + e);
        }
#pragma unroll
        for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
          for (int b_ = 0; b_ < BS; ++b_) A[a_][b_] = fmaf(la[a_], lk[b_], A[a_][b_]);
      }
      if (bi - bj == m + 1) {  // all terms in: L^-1[bi][bj] = -L[bi][bi]^-1 S, in place (descending row index)
        const float* W = Dv + bi * ST;
        float* li = Li + ls_blk(bi, bj) * ST;
#pragma unroll
        for (int a_ = BS - 1; a_ >= 0; --a_) {
          float wr[BS];
#pragma unroll
          for (int k = 0; k < BS; ++k) wr[k] = W[a_ * BS + k];
#pragma unroll
          for (int b_ = 0; b_ < BS; ++b_) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k <= a_; ++k) s = fmaf(wr[k], A[k][b_], s);
            A[a_][b_] = -s;
          }
        }
#pragma unroll
        for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
          for (int b_ = 0; b_ < BS; ++b_) li[a_ * BS + b_] = A[a_][b_];
      }
    }
    sync();
  }

  // ---- G^-1 = L^-T L^-1, block (bi, bj) = sum_{kb >= bi} Li[kb][bi]^T Li[kb][bj]
#pragma unroll
  for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
    for (int b_ = 0; b_ < BS; ++b_) A[a_][b_] = 0.f;
  if (has) {
    for (int kb = bi; kb < NB; ++kb) {
      const float* Pa = Li + ls_blk(kb, bi) * ST;
      const float* Pb = Li + ls_blk(kb, bj) * ST;
#pragma unroll
      for (int k = 0; k < BS; ++k) {
        float pa[BS], pb[BS];
#pragma unroll
        for (int e = 0; e < BS; e += 4) {
          *(float4*)(pa + e) = *(const float4*)(Pa + k * BS + e);
          *(float4*)(pb + e) = *(const float4*)(Pb + k * BS + e);
        }
#pragma unroll
        for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
          for (int b_ = 0; b_ < BS; ++b_) A[a_][b_] = fmaf(pa[a_], pb[b_], A[a_][b_]);
      }
    }
    float* gs = Ls + ls_blk(bi, bj) * ST;  // L is no longer needed: G^-1 blocks, row-major
#pragma unroll
    for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < BS; ++b_) gs[a_ * BS + b_] = A[a_][b_];
  }
  sync();

  // ---- v = G^-1 p (rmhmc/metrics.py:120-127), thread = row
  for (int i = tid; i < D; i += nthr) {
    const int bi_ = i / BS, a_ = i - bi_ * BS;
    float s = 0.f;
    for (int bj_ = 0; bj_ < NB; ++bj_) {
      const bool low = bj_ <= bi_;
      const float* base = low ? Ls + ls_blk(bi_, bj_) * ST + a_ * BS : Ls + ls_blk(bj_, bi_) * ST + a_;
      const int stride = low ? 1 : BS;
#pragma unroll
      for (int k = 0; k < BS; ++k) s = fmaf(base[k * stride], ps[bj_ * BS + k], s);
    }
    vs[i] = s;
    b.v[(size_t)j * D + i] = s;
  }
  for (int i = D + tid; i < DP; i += nthr) vs[i] = 0.f;
  sync();

  // ---- A' = G^-1 - v v^T, packed pairs with the off-diagonal factor 2: x^T A' x = h - u^2 (u = x . v)
  if (has) {
    float* Ap = b.Ap + (size_t)j * d.P;
#pragma unroll
    for (int a_ = 0; a_ < BS; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < BS; ++b_) {
        const int i = BS * bi + a_, jj = BS * bj + b_;
        if (i < D && jj <= i) {
          const float v = fmaf(-vs[i], vs[jj], A[a_][b_]);
          Ap[jj * D - jj * (jj - 1) / 2 + (i - jj)] = (i == jj) ? v : 2.f * v;
        }
      }
  }
  if (tid == 0) {
    float ld = 0.f;
    for (int k = 0; k < NB; ++k) ld += logd[k];
    b.logdet[j] = 2.f * ld;
  }
}

// dH/dq data part and log-density: sums over the m tiles / K tiles of the partials (fixed order: deterministic)
__global__ void __launch_bounds__(256) ls_reduce_kernel(const LsBuf b, const LsDims d) {
  const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
  if (j >= b.n_active[0]) return;
  const int i = blockIdx.y;
  if (i < d.D) {
    float s = 0.f;
    const float* p = b.parts + (size_t)i * d.Ccap + j;
    for (int m = 0; m < d.mtQ; ++m) s += p[(size_t)m * d.D * d.Ccap];
    b.dHt[(size_t)i * d.Ccap + j] = s;
  } else if (b.slot_phase[j] == LS_PH_END) {
    float s = 0.f;
    for (int k = 0; k < d.ktF; ++k) s += b.lp_parts[(size_t)k * d.Ccap + j];
    b.lpt[j] = s;
  }
}

// one warp per running chain: the chain's state machine (see the file header)
__global__ void __launch_bounds__(256) ls_advance_kernel(const LsBuf b, const LsDims d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long j = (long long)blockIdx.x * 8 + warp;
  const long long nact = b.n_active[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) b.n_active[2] = (int)(((long long)b.n_active[2] + nact) & 0x7fffffff);
  if (j >= nact) return;
  const LsCall& call = *b.call;
  const TransArgs& a = call.a;
  const int c = b.idx[j], D = d.D;
  const int ph = b.slot_phase[j];
  const float he = b.he[c], alpha = d.alpha;
  constexpr int EP = 4;  // D <= 128
  float q[EP], p[EP], v[EP], dh[EP], qn[EP], pn[EP];
#pragma unroll
  for (int k = 0; k < EP; ++k) {
    const int i = lane + 32 * k;
    const bool ok = i < D;
    q[k] = ok ? b.q[(size_t)c * D + i] : 0.f;
    p[k] = ok ? b.p[(size_t)c * D + i] : 0.f;
    v[k] = ok ? b.v[(size_t)j * D + i] : 0.f;
    dh[k] = ok ? b.dHt[(size_t)i * d.Ccap + j] : 0.f;
    qn[k] = 0.f;
    pn[k] = 0.f;
    if (ok && ph != LS_PH_END) {
      // rmhmc/integrators.py:119-136: step from the INITIAL coordinates with the gradients at the current guess
      qn[k] = fmaf(he, v[k], b.qi[(size_t)c * D + i]);
      pn[k] = fmaf(-he, fmaf(alpha, q[k], dh[k]), b.pi[(size_t)c * D + i]);  // dH/dq = X^T R + alpha q
    }
  }
  const float half_log_2pi = 0.91893853320467274178f;
  if (ph == LS_PH_FIRST0 || ph == LS_PH_FIRST || ph == LS_PH_ITER) {
    // rmhmc/integrators.py:57-60,64-74: norm = max |x_{n+1} - x_n| over the ravelled (q, p) tuple
    float mx = 0.f;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < EP; ++k) {
      const float dq = fabsf(qn[k] - q[k]), dp = fabsf(pn[k] - p[k]);
      bad = bad || isnan(dq) || isnan(dp);
      mx = fmaxf(mx, fmaxf(dq, dp));
    }
    if (__any_sync(0xffffffffu, bad)) mx = __int_as_float(0x7f800000);
    mx = ls_warp_max(mx);
    const int n_new = (ph == LS_PH_ITER) ? b.nit[c] + 1 : 0;
    const bool cont = (n_new < a.fp_max_iters) && (mx < __int_as_float(0x7f800000)) && (mx < (float)a.fp_div_tol) &&
                      (mx > (float)a.fp_tol);
    if (ph == LS_PH_FIRST0) {
      // initial energy: -logdensity + kinetic (mcmc/metrics.py:160-166, rmhmc/metrics.py:60-74)
      float pw = 0.f;
#pragma unroll
      for (int k = 0; k < EP; ++k) pw = fmaf(p[k], v[k], pw);
      pw = ls_warp_sum(pw);
      if (lane == 0) b.H0[c] = -((const float*)a.out_logp)[c] + 0.5f * pw + 0.5f * b.logdet[j] + half_log_2pi * (float)D;
    }
    if (ph == LS_PH_FIRST0 && a.steps_per_chain != nullptr && a.steps_per_chain[c] <= 0) {
      // dynamic kernel drew zero integration steps: the proposal is the initial state (mcmc/trajectory.py:137)
      if (lane == 0) b.phase[c] = LS_PH_END;
      return;
    }
#pragma unroll
    for (int k = 0; k < EP; ++k) {
      const int i = lane + 32 * k;
      if (i < D) {
        b.q[(size_t)c * D + i] = qn[k];
        b.p[(size_t)c * D + i] = pn[k];
        if (!cont) {  // the explicit update starts from the midpoint :147-148
          b.qi[(size_t)c * D + i] = qn[k];
          b.pi[(size_t)c * D + i] = pn[k];
        } else if (ph != LS_PH_ITER) {  // x0 of this step
          b.qi[(size_t)c * D + i] = q[k];
          b.pi[(size_t)c * D + i] = p[k];
        }
      }
    }
    if (lane == 0) {
      b.nit[c] = n_new;
      if (!cont) b.iters[c] += n_new;
      b.phase[c] = cont ? LS_PH_ITER : LS_PH_EXPL;
    }
  } else if (ph == LS_PH_EXPL) {
#pragma unroll
    for (int k = 0; k < EP; ++k) {
      const int i = lane + 32 * k;
      if (i < D) {
        b.q[(size_t)c * D + i] = qn[k];
        b.p[(size_t)c * D + i] = pn[k];
        b.qi[(size_t)c * D + i] = qn[k];
        b.pi[(size_t)c * D + i] = pn[k];
      }
    }
    if (lane == 0) {
      const int s = b.step[c] + 1;
      b.step[c] = s;
      const int Lc = a.steps_per_chain != nullptr ? a.steps_per_chain[c] : a.num_steps;  // dynamic kernels: per chain
      b.phase[c] = s < Lc ? LS_PH_FIRST : LS_PH_END;
    }
  } else {  // LS_PH_END: energy of the proposal, accept, outputs (rmhmc/rmhmc.py:416-438, mcmc/proposal.py)
    float qq = 0.f, pw = 0.f;
#pragma unroll
    for (int k = 0; k < EP; ++k) {
      qq = fmaf(q[k], q[k], qq);
      pw = fmaf(p[k], v[k], pw);
    }
    qq = ls_warp_sum(qq);
    pw = ls_warp_sum(pw);
    const float lp = b.lpt[j] - 0.5f * alpha * qq;
    const float H1 = -lp + 0.5f * pw + 0.5f * b.logdet[j] + half_log_2pi * (float)D;
    const long long tr = b.tcur[c];
    const float H0 = b.H0[c];
    int acc = 0;
    if (lane == 0) {
      const U2 key = transition_key(a, c, a.ks.first_transition + tr);
      U2 k_m, k_a;
      split2(a.mode, key, k_m, k_a);
      const MH<float> mh = metropolis<float>(a, k_a, c, H0, H1);
      acc = mh.accept;
      store_scalar<float>(a.info.acceptance_rate, c, mh.p_accept);
      if (a.info.is_accepted) a.info.is_accepted[c] = mh.accept;
      if (a.info.is_divergent) a.info.is_divergent[c] = mh.divergent;
      store_scalar<float>(a.info.energy, c, H1);
      store_scalar<float>(a.info.proposal_logdensity, c, lp);
      store_scalar<float>(a.info.proposal_weight, c, mh.weight);
      store_scalar<float>(a.info.initial_energy, c, H0);
      store_scalar<float>(a.info.accept_uniform, c, mh.u);
      if (a.info.fp_iters) a.info.fp_iters[c] = b.iters[c];
      if (a.opts.sample_accept != nullptr) ((float*)a.opts.sample_accept)[tr * a.C + c] = mh.p_accept;
      if (a.opts.accept_sum != nullptr) ((float*)a.opts.accept_sum)[c] += mh.p_accept;
      if (a.opts.dual_averaging != nullptr)
        dual_averaging_update<float>((float*)a.opts.dual_averaging + (size_t)c * 5, mh.p_accept, (float)a.opts.da_target,
                                     (float)a.opts.da_t0, (float)a.opts.da_gamma, (float)a.opts.da_kappa);
      if (mh.accept) ((float*)a.out_logp)[c] = lp;
    }
    acc = __shfl_sync(0xffffffffu, acc, 0);
#pragma unroll
    for (int k = 0; k < EP; ++k) {
      const int i = lane + 32 * k;
      if (i < D) {
        const size_t o = (size_t)c * D + i;
        const float g = -fmaf(alpha, q[k], dh[k]);  // at END R = -(y - s): grad = X^T (y - s) - alpha q
        if (a.info.momentum) ((float*)a.info.momentum)[o] = b.p0[o];
        if (a.info.proposal_position) ((float*)a.info.proposal_position)[o] = q[k];
        if (a.info.proposal_momentum) ((float*)a.info.proposal_momentum)[o] = -p[k];
        if (a.info.proposal_velocity) ((float*)a.info.proposal_velocity)[o] = -v[k];
        if (a.info.proposal_logdensity_grad) ((float*)a.info.proposal_logdensity_grad)[o] = g;
        if (acc) {
          ((float*)a.out_pos)[o] = q[k];
          ((float*)a.out_grad)[o] = g;
        }
        if (a.opts.samples != nullptr)
          ((float*)a.opts.samples)[((size_t)tr * a.C + c) * D + i] = acc ? q[k] : ((const float*)a.out_pos)[o];
      }
    }
    __syncwarp();
    if (tr + 1 < call.T) {
      if (lane == 0) b.tcur[c] = (int)(tr + 1);
      ls_begin_transition(a, b, D, c, tr + 1, lane);
    } else if (lane == 0) {
      b.tcur[c] = (int)(tr + 1);
      b.phase[c] = LS_PH_DONE;
    }
  }
}

// ---- test surface: one round for explicit inputs (gb200_logreg_lockstep_eval) --------------------------------
__global__ void __launch_bounds__(256) ls_eval_setup_kernel(const LsBuf b, const LsDims d, long long C, int mode, const float* q,
                                                            const float* p, const float* qi, const float* pi, float he) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e == 0) { b.n_active[0] = (int)C; b.n_active[1] = 0; b.n_active[2] = 0; }
  if (e < d.Ccap) {
    const unsigned char ph = e < C ? (mode == 1 ? LS_PH_END : mode == 2 ? LS_PH_FIRST0 : LS_PH_ITER) : LS_PH_DONE;
    b.phase[e] = ph;
    if (e < C) { b.idx[e] = (int)e; b.slot_phase[e] = ph; b.he[e] = he; }
  }
  if (e < C * d.D) {
    b.q[e] = q[e];
    if (mode == 2) b.z[e] = p[e]; else b.p[e] = p[e];
    b.qi[e] = qi ? qi[e] : q[e];
    if (mode != 2) b.pi[e] = pi ? pi[e] : p[e];
  }
}
__global__ void __launch_bounds__(256) ls_eval_out_kernel(const LsBuf b, const LsDims d, long long C, int mode, float* qn, float* pn,
                                                          float* p_out, float* logp, float* grad, float* vel, float* logdet,
                                                          float* dHdq) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= C * d.D) return;
  const long long c = e / d.D;
  const int i = (int)(e - c * d.D);
  const float q = b.q[e], v = b.v[e], dh = b.dHt[(size_t)i * d.Ccap + c];
  const float dH = fmaf(d.alpha, q, dh);
  if (vel) vel[e] = v;
  if (mode == 1) {
    if (grad) grad[e] = -dH;
  } else {
    if (dHdq) dHdq[e] = dH;
    if (qn) qn[e] = fmaf(b.he[c], v, b.qi[e]);
    if (pn) pn[e] = fmaf(-b.he[c], dH, b.pi[e]);
    if (p_out && mode == 2) p_out[e] = b.p0[e];
  }
  if (i == 0) {
    if (logdet) logdet[c] = b.logdet[c];
    if (mode == 1 && logp) {
      float qq = 0.f;
      for (int k = 0; k < d.D; ++k) qq = fmaf(b.q[c * d.D + k], b.q[c * d.D + k], qq);
      logp[c] = b.lpt[c] - 0.5f * d.alpha * qq;
    }
  }
}

}  // namespace gb

using namespace gb;

// ---- host side: plan = workspace carve-up + the captured graph -------------------------------------------------
struct gb200_plan {
  gb200_target_desc t;
  LsDims d;
  LsBuf b;
  int64_t ws_bytes;
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaGraphConditionalHandle handle;
  int use_graph;
  int device;
  char graph_note[160];
};

namespace {

int64_t ls_align(int64_t x) { return (x + 255) / 256 * 256; }

int ls_dims(const gb200_target_desc* t, int64_t C, LsDims* d) {
  if (!t || t->kind != GB200_TARGET_LOGREG || t->metric != GB200_METRIC_TARGET) {
    set_error("lock-step plan: needs the logistic-regression target with its Fisher metric");
    return GB200_ERR_INVALID_ARGUMENT;
  }
  if (!t->vec0 || !t->y || t->N < 1 || C < 1) { set_error("lock-step plan: needs vec0 = X^T [D, ldx] (ldx = params[1]), y [N], C >= 1"); return GB200_ERR_INVALID_ARGUMENT; }
  if (t->D < 1 || t->D > 124) { set_error("lock-step plan: D=%d outside 1..124", t->D); return GB200_ERR_UNSUPPORTED; }
  if (C > 0x7fffffffLL / 128) { set_error("lock-step plan: too many chains"); return GB200_ERR_UNSUPPORTED; }
  d->N = (int)t->N;
  d->D = t->D;
  d->ldx = (int)t->params[1];
  if (d->ldx < d->N || d->ldx % 4 != 0) { set_error("lock-step plan: ldx must be >= N and a multiple of 4"); return GB200_ERR_INVALID_ARGUMENT; }
  d->ldn = (int)((C + 3) / 4 * 4);  // row stride of sT[N][ldn] (16-byte rows for the bulk copies)
  d->P = d->D * (d->D + 1) / 2;
  d->PS = ft_ps(d->D);
  d->ktF = (d->N + FT_KT - 1) / FT_KT;
  d->ktQ = d->PS / FT_KT;
  d->mtQ = (d->N + FT_M - 1) / FT_M;
  d->BS = d->D <= 32 ? 4 : 8;
  d->NB = (d->D + d->BS - 1) / d->BS;
  d->nblk = d->NB * (d->NB + 1) / 2;
  const int need = d->nblk > d->D ? d->nblk : d->D;
  d->fthreads = (need + 31) / 32 * 32;
  d->fwarp = (d->BS == 4 && d->nblk <= 32 && d->D <= 32) ? 1 : 0;  // one warp per chain, 4 chains per CTA
  {
    const int ST = d->BS * d->BS + 4;
    d->fsmem_floats = ((2 * d->nblk + d->NB) * ST + 2 * d->NB * d->BS + d->NB + 8 + 3) / 4 * 4;
  }
  d->Ccap = C;
  d->ctiles = (C + FT_N - 1) / FT_N;
  d->ksplit = ft_pick_ksplit((long long)((d->P + FT_M - 1) / FT_M) * d->ctiles, d->ktF, 148);
  d->alpha = (float)t->params[0];
  return GB200_OK;
}

size_t ls_factor_smem(const LsDims& d) { return sizeof(float) * (size_t)d.fsmem_floats * (d.fwarp ? 4 : 1); }

// carve the workspace; with base == NULL only the size is computed
int64_t ls_carve(const LsDims& d, unsigned char* base, LsBuf* b) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += ls_align(bytes);
    return p;
  };
  const int64_t C = d.Ccap, D = d.D;
  LsBuf t;
  memset(&t, 0, sizeof(t));
  t.n_active = (int*)take(64);
  t.call = (LsCall*)take(sizeof(LsCall));
  t.q = (float*)take(C * D * 4); t.p = (float*)take(C * D * 4); t.qi = (float*)take(C * D * 4);
  t.pi = (float*)take(C * D * 4); t.z = (float*)take(C * D * 4); t.p0 = (float*)take(C * D * 4);
  t.he = (float*)take(C * 4); t.H0 = (float*)take(C * 4);
  t.phase = (unsigned char*)take(C);
  t.step = (int*)take(C * 4); t.nit = (int*)take(C * 4); t.tcur = (int*)take(C * 4); t.iters = (int*)take(C * 4);
  t.idx = (int*)take(C * 4);
  t.slot_phase = (unsigned char*)take(C);
  t.v = (float*)take(C * D * 4); t.logdet = (float*)take(C * 4);
  t.sbuf = (float*)take((int64_t)d.N * d.ldn * 4);
  t.Gp = (float*)take(d.ksplit * C * (int64_t)d.PS * 4); t.Ap = (float*)take(C * (int64_t)d.P * 4);
  t.parts = (float*)take((int64_t)d.mtQ * D * C * 4);
  t.lp_parts = (float*)take((int64_t)d.ktF * C * 4);
  t.dHt = (float*)take(D * C * 4); t.lpt = (float*)take(C * 4); t.qT = (float*)take(D * C * 4);
  t.Wt = (unsigned char*)take(d.ctiles * d.ktF * 2 * (int64_t)FT_B_BYTES);
  t.Bt = (unsigned char*)take(d.ctiles * d.ktQ * 2 * (int64_t)FT_B_BYTES);
  t.Xtile = (float*)take((int64_t)d.ktF * D * FT_XS * 4);
  t.pairs = (short2*)take((int64_t)d.PS * sizeof(short2));
  if (b) *b = t;
  return off;
}

// the kernels of one round up to (excluding) the state machine: shared by the sampler loop and the test surface
int ls_launch_eval(const gb200_plan* pl, cudaStream_t s) {
  const LsDims& d = pl->d;
  const LsBuf& b = pl->b;
  {
    dim3 gg((unsigned)((d.Ccap + 31) / 32), (unsigned)((d.D + 31) / 32));
    ls_gather_kernel<<<gg, 256, 0, s>>>(b, d);
    GB_CHECK_LAUNCH();
    ls_weights_kernel<<<(unsigned)(d.ctiles * d.ktF), LS_WTHREADS, sizeof(float) * d.D * FT_KT, s>>>(b, d);
    GB_CHECK_LAUNCH();
  }
  FtArgs a;
  memset(&a, 0, sizeof(a));
  a.Xtile = b.Xtile; a.N = d.N; a.D = d.D; a.Wt = b.Wt; a.C = d.Ccap; a.n_active = b.n_active; a.alpha = d.alpha;
  a.out = b.Gp; a.packed = d.PS; a.ksplit = d.ksplit; a.split_stride = d.Ccap * (long long)d.PS;
  int rc = ft_launch_metric_gemm(a, d.ctiles, s);
  if (rc) return rc;
  if (d.BS == 8 && d.fthreads <= 128) ls_factor_kernel<8, false, 128, 4><<<(unsigned)d.Ccap, d.fthreads, ls_factor_smem(d), s>>>(b, d);
  else if (d.BS == 8) ls_factor_kernel<8, false><<<(unsigned)d.Ccap, d.fthreads, ls_factor_smem(d), s>>>(b, d);
  else if (d.fwarp) ls_factor_kernel<4, true, 128, 8><<<(unsigned)((d.Ccap + 3) / 4), 128, ls_factor_smem(d), s>>>(b, d);
  else ls_factor_kernel<4, false><<<(unsigned)d.Ccap, d.fthreads, ls_factor_smem(d), s>>>(b, d);
  GB_CHECK_LAUNCH();
  rc = ft_launch_quad_b_packed(b.Ap, d.D, d.Ccap, b.n_active, b.Bt, d.ctiles, s);
  if (rc) return rc;
  memset(&a, 0, sizeof(a));
  a.N = d.N; a.D = d.D; a.Wt = b.Bt; a.C = d.Ccap; a.n_active = b.n_active; a.Xt = b.Xt; a.ldx = d.ldx; a.pairs = b.pairs;
  a.PS = d.PS; a.sbuf = b.sbuf; a.lds = d.ldn; a.y = b.y; a.slot_phase = b.slot_phase; a.parts = b.parts; a.Ccap = d.Ccap;
  rc = ft_launch_quad_gemm(a, d.ctiles, 1, s);
  if (rc) return rc;
  dim3 rg((unsigned)((d.Ccap + 255) / 256), (unsigned)(d.D + 1));
  ls_reduce_kernel<<<rg, 256, 0, s>>>(b, d);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

int ls_launch_round(const gb200_plan* pl, cudaStream_t s, int use_handle) {
  int rc = ls_launch_eval(pl, s);
  if (rc) return rc;
  ls_advance_kernel<<<(unsigned)((pl->d.Ccap + 7) / 8), 256, 0, s>>>(pl->b, pl->d);
  GB_CHECK_LAUNCH();
  ls_compact_kernel<<<1, 1024, 0, s>>>(pl->b, pl->d, pl->handle, use_handle, 1);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

#define LS_CUDA(x)                                                            \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      snprintf(pl->graph_note, sizeof(pl->graph_note), "%s: %s", #x, cudaGetErrorString(e_)); \
      goto fail;                                                              \
    }                                                                         \
  } while (0)

// the round as the body of a WHILE node; on any failure the plan falls back to the host-sequenced loop
void ls_build_graph(gb200_plan* pl) {
  pl->use_graph = 0;
  pl->graph = nullptr;
  pl->exec = nullptr;
  bool capturing = false;
  // capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    snprintf(pl->graph_note, sizeof(pl->graph_note), "host-sequenced loop (no capture stream)");
    return;
  }
  cudaGraph_t body = nullptr;
  cudaGraphNode_t node;
  cudaGraphNodeParams np = {};
  LS_CUDA(cudaGraphCreate(&pl->graph, 0));
  LS_CUDA(cudaGraphConditionalHandleCreate(&pl->handle, pl->graph, 1, cudaGraphCondAssignDefault));
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = pl->handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  LS_CUDA(cudaGraphAddNode(&node, pl->graph, nullptr, 0, &np));
  body = np.conditional.phGraph_out[0];
  LS_CUDA(cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  capturing = true;
  if (ls_launch_round(pl, s, 1) != GB200_OK) {
    snprintf(pl->graph_note, sizeof(pl->graph_note), "capture: %s", gb200_last_error());
    goto fail;
  }
  capturing = false;
  LS_CUDA(cudaStreamEndCapture(s, nullptr));
  LS_CUDA(cudaGraphInstantiate(&pl->exec, pl->graph, 0));
  pl->use_graph = 1;
  snprintf(pl->graph_note, sizeof(pl->graph_note), "cuda graph WHILE node");
  cudaStreamDestroy(s);
  return;
fail:
  if (capturing) {
    cudaGraph_t dummy = nullptr;
    cudaStreamEndCapture(s, &dummy);
  }
  cudaGetLastError();
  if (pl->exec) cudaGraphExecDestroy(pl->exec);
  if (pl->graph) cudaGraphDestroy(pl->graph);
  pl->exec = nullptr;
  pl->graph = nullptr;
  pl->use_graph = 0;
  cudaStreamDestroy(s);
  {
    char why[120];
    snprintf(why, sizeof(why), "%s", pl->graph_note);
    snprintf(pl->graph_note, sizeof(pl->graph_note), "host-sequenced loop (%s)", why);
  }
}

}  // namespace

namespace gb {
// gb200_step for rmhmc on the logistic-regression target with gb200_run_opts.plan set (api.cu)
int launch_rmhmc_lockstep(const TransArgs& a, const gb200_target_desc& t, gb200_plan* pl, int dtype, cudaStream_t s) {
  if (dtype != GB200_F32) { set_error("logreg: float32 only"); return GB200_ERR_UNSUPPORTED; }
  if (t.vec0 != pl->t.vec0 || t.y != pl->t.y || t.N != pl->t.N || t.D != pl->t.D || t.params[0] != pl->t.params[0] ||
      t.params[1] != pl->t.params[1] || t.metric != GB200_METRIC_TARGET) {
    set_error("lock-step plan was created for a different target");
    return GB200_ERR_INVALID_ARGUMENT;
  }
  if (a.C > pl->d.Ccap) { set_error("lock-step plan holds %lld chains, launch has %lld", pl->d.Ccap, a.C); return GB200_ERR_INVALID_ARGUMENT; }
  LsCall call;
  memset(&call, 0, sizeof(call));
  call.a = a;
  call.T = a.ks.keys ? 1 : a.ks.num_transitions;
  call.max_rounds = call.T * ((long long)a.num_steps * (a.fp_max_iters + 2) + 1) + 2;
  if (a.num_steps < 1) { set_error("rmhmc (lock-step): num_integration_steps must be >= 1"); return GB200_ERR_INVALID_ARGUMENT; }
  ls_start_kernel<<<(unsigned)((pl->d.Ccap + 7) / 8), 256, 0, s>>>(call, pl->b, pl->d);
  GB_CHECK_LAUNCH();
  ls_compact_kernel<<<1, 1024, 0, s>>>(pl->b, pl->d, pl->handle, 0, 0);
  GB_CHECK_LAUNCH();
  if (pl->use_graph) {
    cudaError_t e = cudaGraphLaunch(pl->exec, s);
    if (e != cudaSuccess) { set_error("lock-step graph launch: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
    count_launch();
    return GB200_OK;
  }
  // host-sequenced fallback (graph WHILE nodes unavailable): poll the running-chain count every few rounds
  for (long long r = 0; r < call.max_rounds;) {
    for (int k = 0; k < 4 && r < call.max_rounds; ++k, ++r) {
      int rc = ls_launch_round(pl, s, 0);
      if (rc) return rc;
    }
    int na = 0;
    cudaError_t e = cudaMemcpyAsync(&na, pl->b.n_active, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("lock-step loop: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
    if (na == 0) break;
  }
  return GB200_OK;
}
}  // namespace gb

extern "C" {

int64_t gb200_rmhmc_logreg_plan_workspace(const gb200_target_desc* t, int64_t C) {
  LsDims d;
  if (ls_dims(t, C, &d)) return -1;
  return ls_carve(d, nullptr, nullptr) + 256;
}

int gb200_rmhmc_logreg_plan_create(const gb200_target_desc* t, int64_t C, void* workspace, int64_t workspace_bytes,
                                   int32_t loop_mode, void* stream, gb200_plan** out) {
  if (!out) { set_error("plan_create: out is NULL"); return GB200_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  LsDims d;
  int rc = ls_dims(t, C, &d);
  if (rc) return rc;
  const int64_t need = ls_carve(d, nullptr, nullptr);
  if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < need) {
    set_error("plan_create: workspace must be 256-byte aligned and >= %lld bytes (gb200_rmhmc_logreg_plan_workspace)", (long long)need);
    return GB200_ERR_INVALID_ARGUMENT;
  }
  if (ls_factor_smem(d) > 227 * 1024 || ft_quad_smem(d.D) > 227 * 1024 || ft_metric_smem(d.D) > 227 * 1024) {
    set_error("plan_create: D=%d does not fit shared memory", d.D);
    return GB200_ERR_UNSUPPORTED;
  }
  gb200_plan* pl = new (std::nothrow) gb200_plan;
  if (!pl) { set_error("plan_create: out of host memory"); return GB200_ERR_CUDA; }
  memset(pl, 0, sizeof(*pl));
  pl->t = *t;
  pl->d = d;
  pl->ws_bytes = workspace_bytes;
  ls_carve(d, (unsigned char*)workspace, &pl->b);
  pl->b.Xt = (const float*)t->vec0;
  pl->b.y = (const float*)t->y;
  cudaGetDevice(&pl->device);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  if (d.BS == 8 && d.fthreads <= 128) e = cudaFuncSetAttribute(ls_factor_kernel<8, false, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ls_factor_smem(d));
  else if (d.BS == 8) e = cudaFuncSetAttribute(ls_factor_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ls_factor_smem(d));
  else if (d.fwarp) e = cudaFuncSetAttribute(ls_factor_kernel<4, true, 128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ls_factor_smem(d));
  else e = cudaFuncSetAttribute(ls_factor_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ls_factor_smem(d));
  if (e != cudaSuccess) { set_error("plan_create: %s", cudaGetErrorString(e)); delete pl; return GB200_ERR_CUDA; }
  rc = ft_set_attributes(d.D);
  if (rc == GB200_OK) rc = ft_launch_xtile(pl->b.Xt, d.ldx, d.N, d.D, pl->b.Xtile, s);
  if (rc == GB200_OK) rc = ft_launch_pairs(pl->b.pairs, d.D, s);
  if (rc) { delete pl; return rc; }
  if (loop_mode == 1) {
    snprintf(pl->graph_note, sizeof(pl->graph_note), "host-sequenced loop (requested)");
  } else {
    ls_build_graph(pl);
  }
  *out = pl;
  return GB200_OK;
}

int gb200_plan_destroy(gb200_plan* pl) {
  if (!pl) return GB200_OK;
  if (pl->exec) cudaGraphExecDestroy(pl->exec);
  if (pl->graph) cudaGraphDestroy(pl->graph);
  delete pl;
  return GB200_OK;
}

const char* gb200_plan_loop_mode(const gb200_plan* pl) { return pl ? pl->graph_note : ""; }

// rounds and chain-evaluations executed by the plan's last launch (synchronises the stream: measurement only)
int gb200_plan_stats(const gb200_plan* pl, int64_t* rounds, int64_t* chain_evals, void* stream) {
  if (!pl) { set_error("plan_stats: plan is NULL"); return GB200_ERR_INVALID_ARGUMENT; }
  int h[3] = {0, 0, 0};
  cudaError_t e = cudaMemcpyAsync(h, pl->b.n_active, 12, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("plan_stats: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  if (rounds) *rounds = h[1];
  if (chain_evals) *chain_evals = h[2];
  return GB200_OK;
}

int gb200_logreg_lockstep_eval(gb200_plan* pl, int32_t mode, const void* q, const void* p, const void* qi, const void* pi,
                               double half_step, void* qn, void* pn, void* p_out, void* logdensity, void* logdensity_grad,
                               void* velocity, void* logdet, void* dHdq, int64_t C, void* stream) {
  if (!pl || !q || !p || mode < 0 || mode > 2 || C < 1 || C > pl->d.Ccap) { set_error("lockstep_eval: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  cudaStream_t s = (cudaStream_t)stream;
  const LsDims& d = pl->d;
  const long long tot = d.Ccap > C * d.D ? d.Ccap : C * d.D;
  ls_eval_setup_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(pl->b, d, C, mode, (const float*)q, (const float*)p,
                                                                    (const float*)qi, (const float*)pi, (float)half_step);
  GB_CHECK_LAUNCH();
  int rc = ls_launch_eval(pl, s);
  if (rc) return rc;
  ls_eval_out_kernel<<<(unsigned)((C * d.D + 255) / 256), 256, 0, s>>>(pl->b, d, C, mode, (float*)qn, (float*)pn, (float*)p_out,
                                                                     (float*)logdensity, (float*)logdensity_grad,
                                                                     (float*)velocity, (float*)logdet, (float*)dHdq);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

}  // extern "C"
