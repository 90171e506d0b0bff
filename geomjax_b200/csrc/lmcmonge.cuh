// Fused lmcmonge transition: Monge-patch metric G = M + alpha^2 grad grad^T handled in closed
// form (Sherman-Morrison), L Lan-integrator steps with the whole chain state in registers.
// Reference: lmcmonge/lmc.py:151-235,512-565; lmcmonge/integrators.py:52-230;
//            lmcmonge/metrics.py:104-284.
#pragma once
#include "transition.cuh"

namespace gb {

// Vectors a lmcmonge integrator state carries (RiemannianIntegratorState, integrators.py:26-44).
// UNIT: inverse_mass_matrix == ones, so dl_ig == dl and ig_Hdl_ig == Hdl_ig (aliases, no storage).
template <typename R, int EPL, bool UNIT>
struct MongeVecs {
  R v[EPL], dl[EPL], Hdl_ig[EPL], Hv[EPL];
  R dl_ig_[UNIT ? 1 : EPL], ig_Hdl_ig_[UNIT ? 1 : EPL], im_[UNIT ? 1 : EPL];
  __device__ __forceinline__ R dl_ig(int k) const { return UNIT ? dl[k] : dl_ig_[k]; }
  __device__ __forceinline__ R ig_Hdl_ig(int k) const { return UNIT ? Hdl_ig[k] : ig_Hdl_ig_[k]; }
  __device__ __forceinline__ R im(int k) const { return UNIT ? R(1) : im_[k]; }
};

// lmcmonge/integrators.py:158-194 (HS=omega as written / omega_fixed) and :197-230 (omegatilde).
// All vectors are distributed; three (two for omegatilde) group reductions.
template <typename R, int EPL, int LPC, bool UNIT>
__device__ __forceinline__ void monge_half_step(int HS, R a2, MongeVecs<R, EPL, UNIT>& m, R& J, R L, R sL, R rs, R eps) {
  const R he = R(0.5) * eps;
  R det1;
  if (HS == GB200_HALF_STEP_OMEGATILDE) {
    Acc4<R, EPL> s0, s1;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      s0.fma(k, m.Hv[k], m.dl_ig(k));
      s1.fma(k, m.dl[k], m.v[k]);
    }
    R p[2] = {s0.total(), s1.total()};
    group_sum_n<LPC>(p);
    det1 = R(1) + he * a2 * p[0];
    const R c = a2 * L * p[1] + he * sL;
    const R c2 = R(0.5) * a2 * eps;
    Acc4<R, EPL> r0, r1;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      m.v[k] += c * m.dl_ig(k) - c2 * m.ig_Hdl_ig(k);
      r0.fma(k, m.dl[k], m.v[k]);
      r1.fma(k, m.Hv[k], m.v[k]);
    }
    R r[2] = {r0.total(), r1.total()};
    group_sum_n<LPC>(r);
    const R f = a2 * (r[0] + he * r[1]) * fast_rcp(det1);
#pragma unroll
    for (int k = 0; k < EPL; ++k) m.v[k] -= f * m.dl_ig(k);
  } else {
    const R a2_sL = a2 * rs;
    R dphi_ig[EPL];
    Acc4<R, EPL> s0, s1;  // Hv.dl_ig, dphi.dl_ig
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const R dphi = fma(a2_sL, m.Hdl_ig[k], -m.dl[k]);
      dphi_ig[k] = UNIT ? dphi : fma(a2_sL, m.ig_Hdl_ig(k), -m.dl_ig(k));
      s0.fma(k, m.Hv[k], m.dl_ig(k));
      s1.fma(k, dphi, m.dl_ig(k));
    }
    R p[2] = {s0.total(), s1.total()};
    group_sum_n<LPC>(p);
    det1 = R(1) + he * a2 * p[0];
    const R hs = he * sL;
    const R ab = a2 * p[1];
    Acc4<R, EPL> s3;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      m.v[k] -= hs * fma(-ab, m.dl_ig(k), dphi_ig[k]);
      s3.fma(k, m.v[k], m.Hv[k]);
    }
    const R d3 = group_sum<LPC>(s3.total());
    R f = he * d3 * fast_rcp(det1);
    if (HS == GB200_HALF_STEP_OMEGA_FIXED) f *= a2;
#pragma unroll
    for (int k = 0; k < EPL; ++k) m.v[k] -= f * m.dl_ig(k);
  }
  Acc4<R, EPL> s4;
#pragma unroll
  for (int k = 0; k < EPL; ++k) s4.fma(k, m.Hdl_ig[k], m.v[k]);
  const R d4 = group_sum<LPC>(s4.total());
  // J += log|1 - he a2 d4| - log|det1|
  J += fast_logabs(R(1) - he * a2 * d4) - fast_logabs(det1);
}

// lmcmonge/metrics.py:168-186 kinetic_energy (mass = 1 / inv_mass elementwise)
template <typename R, int EPL, int LPC, bool UNIT>
__device__ __forceinline__ R monge_kinetic(R a2, const MongeVecs<R, EPL, UNIT>& m, R L, R sum_log_mass) {
  Acc4<R, EPL> s0, s1;
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    s0.fma(k, m.v[k], UNIT ? m.v[k] : m.v[k] / m.im(k));
    s1.fma(k, m.v[k], m.dl[k]);
  }
  R p[2] = {s0.total(), s1.total()};
  group_sum_n<LPC>(p);
  return R(-0.5) * (log(L) + sum_log_mass) + R(0.5) * p[0] + R(0.5) * L * a2 * p[1] * p[1];
}

// velocity_generator, lmcmonge/metrics.py:155-166: v = chol(diag(im) - a2 u u^T) z with u = dl_ig.
// The Cholesky factor of a diagonal-minus-rank-one matrix is L_jj = sqrt(d_j - u_j^2 s_j),
// L_ij = u_i c_j (i > j), c_j = -u_j s_j / L_jj, s_0 = a2, s_{j+1} = s_j d_j / L_jj^2, so
// v_i = L_ii z_i + u_i sum_{k<i} c_k z_k is an O(D) forward recurrence over the elements
// (one rsqrt per element, no division).
template <typename R, class LAY, bool UNIT>
__device__ __forceinline__ void monge_draw(const LAY& lay, R a2, MongeVecs<R, LAY::EPL, UNIT>& m,
                                           const R (&z)[LAY::EPL]) {
  // The recurrence for s is a prefix sum in disguise: 1/s_{j+1} = 1/s_j - u_j^2/d_j, so
  // t_j := 1/s_j = 1/a2 - sum_{k<j} u_k^2/d_k and acc_j = sum_{k<j} c_k z_k are two exclusive
  // scans over the element order j = g + LPC*k; every rsqrt / rcp is then independent (ILP)
  // instead of a D-long chain of dependent MUFU ops.  a2 == 0 gives t = inf, s = 0: v = sqrt(d) z.
  constexpr int EPL = LAY::EPL, LPC = LAY::LPC;
  R run = R(1) / a2;  // t at the start of the current slot row
  R sj[EPL];
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const R u = m.dl_ig(k);
    const R w = lay.valid(k) ? (UNIT ? u * u : u * u / m.im(k)) : R(0);
    R tot;
    const R ex = group_excl_scan<LPC>(w, lay.g, tot);
    sj[k] = fast_rcp(run - ex);
    run -= tot;
  }
  R acc = R(0);
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    const R u = m.dl_ig(k), d = m.im(k);
    const R ljj2 = fma(-u * u, sj[k], d);
    const R r = fast_rsqrt(ljj2);
    const R cz = lay.valid(k) ? -u * sj[k] * r * z[k] : R(0);
    R tot;
    const R ex = group_excl_scan<LPC>(cz, lay.g, tot);
    m.v[k] = lay.valid(k) ? fma(ljj2 * r, z[k], u * (acc + ex)) : R(0);
    acc += tot;
  }
}

// HS >= 0: the half-step variant is a compile-time constant.  LEAN: the launch carries no Info
// outputs, no overrides, no dual averaging / per-chain step size and runs legacy threefry (the
// fused multi-transition launches of a sampling run): all of that code is compiled out.  Both exist
// because the kernel is instruction-cache bound (ncu: 35% of stall samples are no_instruction when
// the hot code exceeds the 32 KB L1.5 I-cache).
// Register cap for the small float layouts (EPL <= 10: c2's (10, 2)): 5 blocks of 128 threads per SM = 96 registers
// (127 uncapped, no spills at 96): 20 instead of 16 resident warps per SM for c2's 27.7 warps per SM.  Measured: no
// change (65.6 vs 65.9 ms per 2048-transition launch) -- the kernel is issue-bound, not occupancy-bound; the same cap
// on lmc's (25, 4) layout changed nothing either (c3: 5.84e9 vs 5.88e9) and was not kept; 72 registers (every warp of c2
// resident in one wave, 76 bytes of spill) was slower: 66.3 vs 62.5 ms.
template <typename R, class Target, int EPL, int LPC, bool EXACT, bool UNIT, int HS = -1, bool LEAN = false>
__global__ void __launch_bounds__(128, (sizeof(R) == 4 && EPL <= 10) ? 5 : 1) lmcmonge_kernel(const TransArgs a, const Target tg) {
  using LAY = Lay<EPL, LPC, EXACT>;
  const int half_step = HS >= 0 ? HS : a.half_step;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long chain = tid / LPC;
  const bool active = chain < a.C;
  if (!active) chain = a.C - 1;  // keep whole warps alive for the shuffles; writes are masked
  LAY lay{a.D, (int)(tid % LPC)};
  const R a2 = (R)a.alpha2;

  MongeVecs<R, EPL, UNIT> m;
  R slm = R(0);  // sum log(mass) = -sum log(inv_mass)
  if (!UNIT) {
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const R imk = (a.inv_mass != nullptr && lay.valid(k))
                        ? ((const R*)a.inv_mass)[chain * a.inv_mass_stride + lay.j(k)] : R(1);
      m.im_[UNIT ? 0 : k] = imk;
      slm -= log(imk);
    }
    slm = group_sum<LPC>(slm);
  }

  const long long T = a.ks.keys ? 1 : a.ks.num_transitions;
  for (long long it = 0; it < T; ++it) {
    const long long t = a.ks.first_transition + it;
    const void* spos = it == 0 ? a.in_pos : a.out_pos;
    const void* slogp = it == 0 ? a.in_logp : a.out_logp;
    const void* sgrad = it == 0 ? a.in_grad : a.out_grad;
    const void* svol = it == 0 ? a.in_vol : a.out_vol;

    R eps = (R)a.step_size;
    R* da = nullptr;
    if (!LEAN) {
      if (a.opts.dual_averaging != nullptr) {
        da = (R*)a.opts.dual_averaging + chain * 5;
        eps = exp(da[0]);
      } else if (a.step_size_per_chain != nullptr) {
        eps = ((const R*)a.step_size_per_chain)[chain];
      }
    }

    R q[EPL];
    load_vec(lay, spos, chain, q);
    load_vec(lay, sgrad, chain, m.dl);  // un-normalised gradient g0 for now
    const R l0 = ((const R*)slogp)[chain];
    const R J0 = ((const R*)svol)[chain];

    // ---- prologue: lmcmonge/lmc.py:177-208
    R L;
    {
      Acc4<R, EPL> s;
#pragma unroll
      for (int k = 0; k < EPL; ++k) s.fma(k, m.im(k) * m.dl[k], m.dl[k]);
      L = R(1) + a2 * group_sum<LPC>(s.total());  // normalizing_constant, metrics.py:193-197
    }
    R rs = fast_rsqrt(L);
    R sL = L * rs;
    const R sL0 = sL;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      m.dl[k] *= rs;
      if (!UNIT) m.dl_ig_[UNIT ? 0 : k] = m.im(k) * m.dl[k];
    }
    U2 k_v, k_a;
    if (LEAN && LPC >= 2) {  // lean = legacy threefry: the two lanes of a pair share the key-tree blocks
      const U2 key = transition_key_shared(a, chain, t, lay.g);
      split2_shared(key, lay.g, k_v, k_a);
    } else {
      const U2 key = transition_key(a, chain, t);
      split2(LEAN ? GB200_THREEFRY_LEGACY : a.mode, key, k_v, k_a);
    }
    {
      R z[EPL];
      draw_noise<R, LAY, LEAN>(a, lay, k_v, chain, z);
      monge_draw<R, LAY, UNIT>(lay, a2, m, z);
      if (!LEAN && active) store_vec(lay, a.info.noise, chain, z);
    }
    if (!LEAN && active) store_vec(lay, a.info.momentum, chain, m.v);  // LMCInfo.velocity = the initial draw
    typename Target::Ctx ctx = tg.prepare(lay, q);
    {
      R u[EPL];
#pragma unroll
      for (int k = 0; k < EPL; ++k) u[k] = m.dl_ig(k);
      tg.hvp2(lay, ctx, q, u, m.v, rs, m.Hdl_ig, m.Hv);
    }
    if (!UNIT) {
#pragma unroll
      for (int k = 0; k < EPL; ++k) m.ig_Hdl_ig_[UNIT ? 0 : k] = m.im(k) * m.Hdl_ig[k];
    }

    R J = J0;
    const R H0 = -l0 + monge_kinetic<R, EPL, LPC, UNIT>(a2, m, L, slm) - J0;  // lmcmonge_energy
    R lp = l0;

    // ---- L integrator steps: lmcmonge/integrators.py:63-153.  One step = half-step, position /
    // gradient / HVP refresh, half-step, Hv refresh; written as 2L half-steps so that the half-step
    // body exists once in the instruction stream (same operations in the same order).
    // Dynamic kernels (lmcmonge/lmc.py build_dynamic_kernel): a per-chain step count.  The warp runs to its largest
    // count; a chain that is done keeps stepping with step size 0, which is an exact no-op of every update below
    // (v - 0 * x, q + 0 * v, log|1 + 0|), so its state is bit-identical to having stopped.
    int nh_chain = 2 * a.num_steps, nh = nh_chain;
    if (!LEAN && a.steps_per_chain != nullptr) {
      nh_chain = 2 * a.steps_per_chain[chain];
      nh = __reduce_max_sync(0xffffffffu, nh_chain);
    }
#pragma unroll 1
    for (int h = 0; h < nh; ++h) {
      const R eps_h = (LEAN || h < nh_chain) ? eps : R(0);
      monge_half_step<R, EPL, LPC, UNIT>(half_step, a2, m, J, L, sL, rs, eps_h);
      if (!(h & 1)) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) q[k] = fma(eps_h, m.v[k], q[k]);
        ctx = tg.prepare(lay, q);
        lp = tg.logp(ctx);
        if constexpr (UNIT && Target::kGradSqnorm) {
          // unit mass: L = 1 + a2 |grad|^2 straight from the context, dl = grad / sqrt(L) in one pass
          L = R(1) + a2 * tg.grad_sqnorm(ctx);
          rs = fast_rsqrt(L);
          sL = L * rs;
          tg.grad_scaled(lay, ctx, q, rs, m.dl);
          R u[EPL];
#pragma unroll
          for (int k = 0; k < EPL; ++k) u[k] = m.dl[k];
          tg.hvp2(lay, ctx, q, u, m.v, rs, m.Hdl_ig, m.Hv);
        } else {
          tg.grad(lay, ctx, q, m.dl);  // un-normalised gradient for now
          {
            Acc4<R, EPL> sg;
#pragma unroll
            for (int k = 0; k < EPL; ++k) sg.fma(k, m.im(k) * m.dl[k], m.dl[k]);
            L = R(1) + a2 * group_sum<LPC>(sg.total());
          }
          rs = fast_rsqrt(L);
          sL = L * rs;
          R u[EPL];
#pragma unroll
          for (int k = 0; k < EPL; ++k) {
            m.dl[k] *= rs;
            if (!UNIT) m.dl_ig_[UNIT ? 0 : k] = m.im(k) * m.dl[k];
            u[k] = m.dl_ig(k);
          }
          tg.hvp2(lay, ctx, q, u, m.v, rs, m.Hdl_ig, m.Hv);
        }
        if (!UNIT) {
#pragma unroll
          for (int k = 0; k < EPL; ++k) m.ig_Hdl_ig_[UNIT ? 0 : k] = m.im(k) * m.Hdl_ig[k];
        }
      } else if (h + 1 < nh) {
        tg.hvp(lay, ctx, q, m.v, rs, m.Hv);  // :132-134 (only feeds the next step)
      }
    }

    // ---- flip, energy, accept: lmcmonge/lmc.py:512-534
    // energy is even in v, so evaluate on v and store -v
    const R H1 = -lp + monge_kinetic<R, EPL, LPC, UNIT>(a2, m, L, slm) - J;
    MH<R> mh = metropolis<R, LEAN>(a, k_a, chain, H0, H1);

    if (!LEAN && a.info.proposal_momentum != nullptr) {
      // metric_vector_product with the un-normalised gradient g = dl * sL (integrators.py:136-138)
      R d = group_sum<LPC>(dotv<R, EPL>(m.v, m.dl)) * sL;
      const R c = a2 * L * d * sL;
      R pm[EPL];
#pragma unroll
      for (int k = 0; k < EPL; ++k) pm[k] = m.v[k] / m.im(k) + c * m.dl[k];
      if (active) store_vec(lay, a.info.proposal_momentum, chain, pm, R(-1));
    }
    R gp[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) gp[k] = m.dl[k] * sL;  // lmc.py:226-233
    if (!LEAN && active) {
      store_vec(lay, a.info.proposal_position, chain, q);
      store_vec(lay, a.info.proposal_velocity, chain, m.v, R(-1));
      store_vec(lay, a.info.proposal_logdensity_grad, chain, gp);
      if (lay.g == 0) {
        store_scalar<R>(a.info.acceptance_rate, chain, mh.p_accept);
        if (a.info.is_accepted) a.info.is_accepted[chain] = mh.accept;
        if (a.info.is_divergent) a.info.is_divergent[chain] = mh.divergent;
        store_scalar<R>(a.info.energy, chain, H1);
        store_scalar<R>(a.info.proposal_logdensity, chain, lp);
        store_scalar<R>(a.info.proposal_volume_adjustment, chain, J);
        store_scalar<R>(a.info.proposal_weight, chain, mh.weight);
        store_scalar<R>(a.info.initial_energy, chain, H0);
        store_scalar<R>(a.info.accept_uniform, chain, mh.u);
      }
    }
    if (!mh.accept) {
      // rejected: keep the input state; the returned gradient is dl0 * sqrt(L0) (lmc.py:226-233)
      load_vec(lay, spos, chain, q);
      load_vec(lay, sgrad, chain, gp);
      const R rs0 = R(1) / sL0;
#pragma unroll
      for (int k = 0; k < EPL; ++k) gp[k] = (gp[k] * rs0) * sL0;
      lp = l0;
      J = J0;
    }
    if (active) {
      store_vec(lay, a.out_pos, chain, q);
      store_vec(lay, a.out_grad, chain, gp);
      if (a.opts.samples != nullptr)
        store_vec(lay, (R*)a.opts.samples + it * a.C * (long long)a.D, chain, q);
      if (lay.g == 0) {
        store_scalar<R>(a.out_logp, chain, lp);
        store_scalar<R>(a.out_vol, chain, J);
        if (a.opts.sample_accept != nullptr) ((R*)a.opts.sample_accept)[it * a.C + chain] = mh.p_accept;
        if (a.opts.accept_sum != nullptr) ((R*)a.opts.accept_sum)[chain] += mh.p_accept;
        if (!LEAN && da != nullptr)
          dual_averaging_update<R>(da, mh.p_accept, (R)a.opts.da_target, (R)a.opts.da_t0, (R)a.opts.da_gamma,
                                   (R)a.opts.da_kappa);
      }
    }
    __syncwarp();  // scalars written by lane 0 are re-read by the whole group next transition
  }
}

}  // namespace gb
