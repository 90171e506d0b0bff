// Pieces shared by the three fused transition kernels: argument block, per-chain key
// derivation, noise draw, Metropolis-Hastings accept, dual-averaging epilogue.
#pragma once
#include "targets.cuh"

namespace gb {

// POD argument block (passed by value; built by api.cu from the C-ABI structs).
struct TransArgs {
  // state in / out (field-wise aliasing allowed)
  const void *in_pos, *in_logp, *in_grad, *in_vol;
  void *out_pos, *out_logp, *out_grad, *out_vol;
  gb200_info info;
  gb200_run_opts opts;
  gb200_key_source ks;
  // kernel params
  double step_size;
  const void* step_size_per_chain;
  const void* inv_mass;
  long long inv_mass_stride;  // 0: shared [D]; D: per chain [C, D]
  double alpha2;
  double divergence_threshold;
  double fp_tol, fp_div_tol;
  int fp_max_iters;
  int num_steps;
  const int* steps_per_chain;  // optional [C]: per-chain number of integrator steps (dynamic kernels); num_steps = their upper bound
  int half_step;
  int D;
  int metric;  // gb200_metric_kind
  int mode;    // gb200_threefry_mode
  long long C;
  // logreg tcgen05 lock-step path: chains whose fixed point needs more than lock_cap iterations are
  // appended to work_list and re-run chain by chain (rmhmc_logreg.cu) instead of stalling their tile
  int* work_count;
  int* work_list;
  int lock_cap;
};

__device__ __forceinline__ U2 transition_key(const TransArgs& a, long long chain, long long t) {
  if (a.ks.keys != nullptr) {
    return U2{a.ks.keys[2 * chain], a.ks.keys[2 * chain + 1]};
  }
  U2 root{a.ks.root_key[0], a.ks.root_key[1]};
  return chain_key(a.mode, root, (uint32_t)a.ks.total_transitions, (uint32_t)t,
                   (uint32_t)a.ks.total_chains, (uint32_t)(a.ks.chain_offset + chain));
}

// Lean launches (legacy threefry) with >= 2 lanes per chain: the key tree costs six threefry blocks per transition
// (two per split) that every lane of a chain would repeat.  Lanes g and g ^ 1 of the chain hash ONE block of each
// split and exchange the word by shuffle: three blocks per lane instead of six, same keys bit for bit.
__device__ __forceinline__ U2 split_index_shared(U2 key, uint32_t num, uint32_t i, int g) {
  const uint32_t j = 2u * i + (uint32_t)(g & 1);  // element of random_bits(key, 2 num) this lane produces
  const bool y = j >= num;
  const U2 o = threefry2x32(key.x, key.y, y ? j - num : j, y ? j : j + num);
  const uint32_t mine = y ? o.y : o.x;
  const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
  return (g & 1) ? U2{other, mine} : U2{mine, other};
}
__device__ __forceinline__ U2 transition_key_shared(const TransArgs& a, long long chain, long long t, int g) {
  if (a.ks.keys != nullptr) return U2{a.ks.keys[2 * chain], a.ks.keys[2 * chain + 1]};
  const U2 kt = split_index_shared(U2{a.ks.root_key[0], a.ks.root_key[1]}, (uint32_t)a.ks.total_transitions, (uint32_t)t, g);
  return split_index_shared(kt, (uint32_t)a.ks.total_chains, (uint32_t)(a.ks.chain_offset + chain), g);
}
// (k_a, k_b) = split(key, 2): bits = hash of counts (0, 2) and (1, 3); lane parity p hashes block (p, p + 2)
__device__ __forceinline__ void split2_shared(U2 key, int g, U2& ka, U2& kb) {
  const uint32_t p = (uint32_t)(g & 1);
  const U2 o = threefry2x32(key.x, key.y, p, p + 2u);  // o.x = bits[p], o.y = bits[p + 2]
  const uint32_t ox = __shfl_xor_sync(0xffffffffu, o.x, 1), oy = __shfl_xor_sync(0xffffffffu, o.y, 1);
  ka = p ? U2{ox, o.x} : U2{o.x, ox};
  kb = p ? U2{oy, o.y} : U2{o.y, oy};
}

// z = jax.random.normal(key, (D,)) distributed over the lane group (util.py:81-82).
// Legacy threefry hashes the counters pairwise (i, i + D/2): when the layout is exact, D is
// even and both halves of a pair live in the same lane, one block yields two normals.
// The generation loop is deliberately NOT unrolled (threefry + erfinv is ~250 SASS instructions
// per element; unrolling it EPL times blew the instruction cache: 45% stall_no_inst at EPL=25).
// Values are staged through this thread's private column of shared memory so that the
// register array z[] is still filled with static indices.
extern __shared__ unsigned char gb_smem[];

// LEAN: the launch has no noise override and runs legacy threefry (checked by the host), so only
// one generation loop is compiled in (code size: the fused kernels are instruction-cache bound).
template <typename R, class LAY, bool LEAN = false>
__device__ __forceinline__ void draw_noise(const TransArgs& a, const LAY& lay, U2 key, long long chain,
                                           R (&z)[LAY::EPL]) {
  constexpr int EPL = LAY::EPL, LPC = LAY::LPC;
  if (!LEAN && a.opts.noise_override != nullptr) {
    const R* zo = (const R*)a.opts.noise_override + chain * a.D;
#pragma unroll
    for (int k = 0; k < EPL; ++k) z[k] = lay.valid(k) ? zo[lay.j(k)] : R(0);
    return;
  }
  R* zs = (R*)gb_smem + threadIdx.x;
  const int stride = blockDim.x;
  constexpr int DS = EPL * LPC;
  constexpr bool PAIRED = LAY::EXACT && (DS % 2 == 0) && ((DS / 2) % LPC == 0);
  constexpr bool CROSS = LAY::EXACT && !PAIRED && LPC > 1 && (DS % 2 == 0) && ((DS / 2) % LPC == LPC / 2);
  const int mode = LEAN ? GB200_THREEFRY_LEGACY : a.mode;
  if (PAIRED && mode == GB200_THREEFRY_LEGACY) {
    constexpr int HK = EPL / 2;  // slots per half
    int k = 0;
#pragma unroll 1
    for (; k + 1 < HK; k += 2) {  // two blocks per call (interleaved rounds), four normals
      const uint32_t j0 = (uint32_t)lay.j(k), j1 = (uint32_t)lay.j(k + 1);
      const U4 o = threefry2x32_x2(key.x, key.y, j0, j0 + (uint32_t)(DS / 2), j1, j1 + (uint32_t)(DS / 2));
      const float2 na = bits_to_normal_x2(o.a0, o.a1);
      const float2 nb = bits_to_normal_x2(o.b0, o.b1);
      zs[k * stride] = (R)na.x;
      zs[(k + HK) * stride] = (R)na.y;
      zs[(k + 1) * stride] = (R)nb.x;
      zs[(k + 1 + HK) * stride] = (R)nb.y;
    }
    if (k < HK) {
      const uint32_t j = (uint32_t)lay.j(k);
      U2 o = threefry2x32(key.x, key.y, j, j + (uint32_t)(DS / 2));
      const float2 nz = bits_to_normal_x2(o.x, o.y);
      zs[k * stride] = (R)nz.x;
      zs[(k + HK) * stride] = (R)nz.y;
    }
  } else if (CROSS && mode == GB200_THREEFRY_LEGACY) {
    // D even but the two halves of a legacy counter pair (j, j + D/2) live in lanes g and g ^ (LPC/2):
    // hash each pair once in the lane that owns j, hand the second word to the partner lane with one
    // shuffle (c3: D = 100 over 4 lanes -> 50 threefry blocks per chain instead of 100).
    constexpr int H = DS / 2;
    constexpr int KH = (H + LPC - 1) / LPC;  // slots whose element index can be < H
    const int gp = lay.g ^ (LPC / 2);
    const int off = (H + gp - lay.g) / LPC;  // slot of element (partner's j) + H in this lane
#pragma unroll 1
    for (int k = 0; k < KH; ++k) {
      const uint32_t j = (uint32_t)lay.j(k);
      const U2 o = threefry2x32(key.x, key.y, j, j + (uint32_t)H);
      const uint32_t y = __shfl_xor_sync(0xffffffffu, o.y, LPC / 2);
      const float2 nz = bits_to_normal_x2(o.x, y);
      if ((int)j < H) zs[k * stride] = (R)nz.x;
      if (gp + LPC * k < H) zs[(k + off) * stride] = (R)nz.y;
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < EPL; ++k) {  // 4 independent threefry + erfinv chains in flight (ILP)
      R v = R(0);
      if (lay.valid(k)) v = (R)bits_to_normal_call(random_bits_elem(mode, key, (uint32_t)lay.j(k), (uint32_t)lay.D()));
      zs[k * stride] = v;
    }
  }
#pragma unroll
  for (int k = 0; k < EPL; ++k) z[k] = zs[k * stride];
}

// mcmc/proposal.py:87-121 (proposal_from_energy_diff) + :168-185 (static_binomial_sampling)
template <typename R>
struct MH {
  R weight, p_accept, u;
  bool accept, divergent;
};

template <typename R, bool LEAN = false>
__device__ __forceinline__ MH<R> metropolis(const TransArgs& a, U2 key_accept, long long chain, R H0, R H1) {
  MH<R> m;
  R delta = H0 - H1;
  if (isnan(delta)) delta = -Lim<R>::inf();
  m.weight = delta;
  m.p_accept = fmin(exp(delta), R(1));
  m.divergent = (-delta) > (R)a.divergence_threshold;
  if (!LEAN && a.opts.uniform_override != nullptr) m.u = ((const R*)a.opts.uniform_override)[chain];
  else m.u = (R)uniform_scalar(LEAN ? GB200_THREEFRY_LEGACY : a.mode, key_accept);
  m.accept = m.u < m.p_accept;
  return m;
}

// optimizers/dual_averaging.py:101-123 applied to gradient = target - acceptance_rate
template <typename R>
__device__ __forceinline__ void dual_averaging_update(R* da, R accept_rate, R target, R t0, R gamma, R kappa) {
  const R log_step = da[0], avg_log_step = da[1], step = da[2], avg_err0 = da[3], mu = da[4];
  const R reg_step = step + t0;
  const R eta = pow(step, -kappa);
  const R avg_err = (R(1) - R(1) / reg_step) * avg_err0 + (target - accept_rate) / reg_step;
  da[0] = mu - (sqrt(step) / gamma) * avg_err;
  da[1] = eta * log_step + (R(1) - eta) * avg_log_step;
  da[2] = step + R(1);
  da[3] = avg_err;
}

template <typename R, class LAY>
__device__ __forceinline__ void load_vec(const LAY& lay, const void* base, long long chain, R (&v)[LAY::EPL]) {
  const R* p = (const R*)base + chain * lay.D();
#pragma unroll
  for (int k = 0; k < LAY::EPL; ++k) v[k] = lay.valid(k) ? p[lay.j(k)] : R(0);
}
template <typename R, class LAY>
__device__ __forceinline__ void store_vec(const LAY& lay, void* base, long long chain, const R (&v)[LAY::EPL],
                                          R sign = R(1)) {
  if (base == nullptr) return;
  R* p = (R*)base + chain * lay.D();
#pragma unroll
  for (int k = 0; k < LAY::EPL; ++k)
    if (lay.valid(k)) p[lay.j(k)] = sign * v[k];
}
template <typename R>
__device__ __forceinline__ void store_scalar(void* base, long long chain, R v) {
  if (base != nullptr) ((R*)base)[chain] = v;
}

// fast scalar helpers (float: MUFU-based intrinsics, <= 2 ulp; double: exact library calls)
__device__ __forceinline__ float fast_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double fast_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float fast_rcp(float x) { return __fdividef(1.0f, x); }
__device__ __forceinline__ double fast_rcp(double x) { return 1.0 / x; }
// log|x| for volume adjustments: |abs error| <= 2^-21.4 for x in [0.5, 2] (CUDA __logf), which is
// below the float32 resolution of the O(10) energies it is added to.
__device__ __forceinline__ float fast_logabs(float x) { return __logf(fabsf(x)); }
__device__ __forceinline__ double fast_logabs(double x) { return log(fabs(x)); }

}  // namespace gb
