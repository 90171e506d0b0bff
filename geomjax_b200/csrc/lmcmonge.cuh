// Fused lmcmonge transition: Monge-patch metric G = M + alpha^2 grad grad^T handled in closed
// form (Sherman-Morrison), L Lan-integrator steps with the whole chain state in registers.
// Reference: lmcmonge/lmc.py:151-235,512-565; lmcmonge/integrators.py:52-230;
//            lmcmonge/metrics.py:104-284.
#pragma once
#include "transition.cuh"

namespace gb {

// lmcmonge/integrators.py:158-194 (HS=omega as written / omega_fixed) and :197-230 (omegatilde).
// All vectors are distributed; three (two for omegatilde) group reductions.
template <typename R, int EPL, int LPC>
__device__ __forceinline__ void monge_half_step(int HS, R a2, R (&v)[EPL], R& J, const R (&dl)[EPL], const R (&Hv)[EPL],
                                                R L, R sL, const R (&dl_ig)[EPL], const R (&Hdl_ig)[EPL],
                                                const R (&ig_Hdl_ig)[EPL], R eps) {
  const R he = R(0.5) * eps;
  if (HS == GB200_HALF_STEP_OMEGATILDE) {
    R p[2] = {dotv<R, EPL, LPC>(Hv, dl_ig), dotv<R, EPL, LPC>(dl, v)};
    group_sum_n<LPC>(p);
    const R det1 = R(1) + he * a2 * p[0];
    J -= log(fabs(det1));
    const R c = a2 * L * p[1] + he * sL;
#pragma unroll
    for (int k = 0; k < EPL; ++k) v[k] += c * dl_ig[k] - R(0.5) * a2 * eps * ig_Hdl_ig[k];
    R r[2] = {dotv<R, EPL, LPC>(dl, v), dotv<R, EPL, LPC>(Hv, v)};
    group_sum_n<LPC>(r);
    const R f = a2 * (r[0] + he * r[1]) / det1;
#pragma unroll
    for (int k = 0; k < EPL; ++k) v[k] -= f * dl_ig[k];
  } else {
    const R a2_sL = a2 / sL;
    R dphi_ig[EPL];
    R p[2] = {R(0), R(0)};  // Hv.dl_ig, dphi.dl_ig
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const R dphi = a2_sL * Hdl_ig[k] - dl[k];
      dphi_ig[k] = a2_sL * ig_Hdl_ig[k] - dl_ig[k];
      p[0] += Hv[k] * dl_ig[k];
      p[1] += dphi * dl_ig[k];
    }
    group_sum_n<LPC>(p);
    const R det1 = R(1) + he * a2 * p[0];
    J -= log(fabs(det1));
    const R hs = he * sL;
    const R ab = a2 * p[1];
    R d3 = R(0);
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      v[k] -= hs * (dphi_ig[k] - ab * dl_ig[k]);
      d3 += v[k] * Hv[k];
    }
    d3 = group_sum<LPC>(d3);
    R f = he * d3 / det1;
    if (HS == GB200_HALF_STEP_OMEGA_FIXED) f *= a2;
#pragma unroll
    for (int k = 0; k < EPL; ++k) v[k] -= f * dl_ig[k];
  }
  R d4 = group_sum<LPC>(dotv<R, EPL, LPC>(Hdl_ig, v));
  J += log(fabs(R(1) - he * a2 * d4));
}

// lmcmonge/metrics.py:168-186 kinetic_energy (mass = 1 / inv_mass elementwise)
template <typename R, int EPL, int LPC>
__device__ __forceinline__ R monge_kinetic(R a2, const R (&v)[EPL], const R (&dl)[EPL], const R (&im)[EPL], R L,
                                           R sum_log_mass) {
  R p[2] = {R(0), R(0)};
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    p[0] += v[k] * v[k] / im[k];
    p[1] += v[k] * dl[k];
  }
  group_sum_n<LPC>(p);
  return R(-0.5) * (log(L) + sum_log_mass) + R(0.5) * p[0] + R(0.5) * L * a2 * p[1] * p[1];
}

// velocity_generator, lmcmonge/metrics.py:155-166: v = chol(diag(im) - a2 u u^T) z with u = dl_ig.
// The Cholesky factor of a diagonal-minus-rank-one matrix is L_jj = sqrt(d_j - u_j^2 s_j),
// L_ij = u_i c_j (i > j), c_j = -u_j s_j / L_jj, s_0 = a2, s_{j+1} = s_j d_j / L_jj^2, so
// v_i = L_ii z_i + u_i sum_{k<i} c_k z_k is an O(D) forward recurrence over the elements.
template <typename R, int EPL, int LPC>
__device__ __forceinline__ void monge_draw(const Lay<EPL, LPC>& lay, R a2, const R (&im)[EPL], const R (&u)[EPL],
                                           const R (&z)[EPL], R (&v)[EPL]) {
  R s = a2, acc = R(0);
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
#pragma unroll(LPC <= 2 ? LPC : 1)
    for (int gg = 0; gg < LPC; ++gg) {
      // every lane evaluates with its own slot-k data; only the owner lane gg is meaningful
      const R ljj2 = im[k] - u[k] * u[k] * s;
      const R ljj = sqrt(ljj2);
      const R cz = -u[k] * s / ljj * z[k];
      const R snext = s * im[k] / ljj2;
      if (lay.g == gg) v[k] = lay.valid(k) ? ljj * z[k] + u[k] * acc : R(0);
      const bool in = (gg + LPC * k) < lay.D;  // uniform across the group
      if (LPC == 1) {
        if (in) { acc += cz; s = snext; }
      } else {
        const R czb = group_bcast<LPC>(cz, gg);
        const R snb = group_bcast<LPC>(snext, gg);
        if (in) { acc += czb; s = snb; }
      }
    }
  }
}

template <typename R, class Target, int EPL, int LPC>
__global__ void __launch_bounds__(128) lmcmonge_kernel(const TransArgs a, const Target tg) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long chain = tid / LPC;
  const bool active = chain < a.C;
  if (!active) chain = a.C - 1;  // keep whole warps alive for the shuffles; writes are masked
  Lay<EPL, LPC> lay{a.D, (int)(tid % LPC)};
  const R a2 = (R)a.alpha2;

  R im[EPL];
  R slm = R(0);  // sum log(mass) = -sum log(inv_mass)
#pragma unroll
  for (int k = 0; k < EPL; ++k) {
    im[k] = (a.inv_mass != nullptr && lay.valid(k)) ? ((const R*)a.inv_mass)[lay.j(k)] : R(1);
    slm -= log(im[k]);
  }
  slm = group_sum<LPC>(slm);

  const long long T = a.ks.keys ? 1 : a.ks.num_transitions;
  for (long long it = 0; it < T; ++it) {
    const long long t = a.ks.first_transition + it;
    const void* spos = it == 0 ? a.in_pos : a.out_pos;
    const void* slogp = it == 0 ? a.in_logp : a.out_logp;
    const void* sgrad = it == 0 ? a.in_grad : a.out_grad;
    const void* svol = it == 0 ? a.in_vol : a.out_vol;

    R eps = (R)a.step_size;
    R* da = nullptr;
    if (a.opts.dual_averaging != nullptr) {
      da = (R*)a.opts.dual_averaging + chain * 5;
      eps = exp(da[0]);
    } else if (a.step_size_per_chain != nullptr) {
      eps = ((const R*)a.step_size_per_chain)[chain];
    }

    R q[EPL], g0[EPL];
    load_vec(lay, spos, chain, q);
    load_vec(lay, sgrad, chain, g0);
    const R l0 = ((const R*)slogp)[chain];
    const R J0 = ((const R*)svol)[chain];

    // ---- prologue: lmcmonge/lmc.py:177-208
    R L;
    {
      R s = R(0);
#pragma unroll
      for (int k = 0; k < EPL; ++k) s += im[k] * g0[k] * g0[k];
      L = R(1) + a2 * group_sum<LPC>(s);  // normalizing_constant, metrics.py:193-197
    }
    R sL = sqrt(L);
    R rs = R(1) / sL;
    const R sL0 = sL, rs0 = rs;
    R dl[EPL], dl_ig[EPL], Hdl_ig[EPL], ig_Hdl_ig[EPL], Hv[EPL], v[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      dl[k] = g0[k] * rs;
      dl_ig[k] = im[k] * dl[k];
    }
    U2 key = transition_key(a, chain, t);
    U2 k_v, k_a;
    split2(a.mode, key, k_v, k_a);
    R z[EPL];
    draw_noise<R>(a, lay, k_v, chain, z);
    monge_draw(lay, a2, im, dl_ig, z, v);
    if (active) {
      store_vec(lay, a.info.noise, chain, z);
      store_vec(lay, a.info.momentum, chain, v);  // LMCInfo.velocity = the initial draw
    }
    typename Target::Ctx ctx = tg.prepare(lay, q);
    tg.hvp2(lay, ctx, q, dl_ig, v, rs, Hdl_ig, Hv);
#pragma unroll
    for (int k = 0; k < EPL; ++k) ig_Hdl_ig[k] = im[k] * Hdl_ig[k];

    R J = J0;
    const R H0 = -l0 + monge_kinetic<R, EPL, LPC>(a2, v, dl, im, L, slm) - J0;  // lmcmonge_energy
    R lp = l0;

    // ---- L integrator steps: lmcmonge/integrators.py:63-153
    for (int s = 0; s < a.num_steps; ++s) {
      monge_half_step<R, EPL, LPC>(a.half_step, a2, v, J, dl, Hv, L, sL, dl_ig, Hdl_ig, ig_Hdl_ig, eps);
#pragma unroll
      for (int k = 0; k < EPL; ++k) q[k] += eps * v[k];
      ctx = tg.prepare(lay, q);
      lp = tg.logp(ctx);
      tg.grad(lay, ctx, q, dl);  // un-normalised gradient for now
      {
        R sg = R(0);
#pragma unroll
        for (int k = 0; k < EPL; ++k) sg += im[k] * dl[k] * dl[k];
        L = R(1) + a2 * group_sum<LPC>(sg);
      }
      sL = sqrt(L);
      rs = R(1) / sL;
#pragma unroll
      for (int k = 0; k < EPL; ++k) {
        dl[k] *= rs;
        dl_ig[k] = im[k] * dl[k];
      }
      tg.hvp2(lay, ctx, q, dl_ig, v, rs, Hdl_ig, Hv);
#pragma unroll
      for (int k = 0; k < EPL; ++k) ig_Hdl_ig[k] = im[k] * Hdl_ig[k];
      monge_half_step<R, EPL, LPC>(a.half_step, a2, v, J, dl, Hv, L, sL, dl_ig, Hdl_ig, ig_Hdl_ig, eps);
      if (s + 1 < a.num_steps) tg.hvp(lay, ctx, q, v, rs, Hv);  // :132-134 (only feeds the next step)
    }

    // ---- flip, energy, accept: lmcmonge/lmc.py:512-534
    // energy is even in v, so evaluate on v and store -v
    const R H1 = -lp + monge_kinetic<R, EPL, LPC>(a2, v, dl, im, L, slm) - J;
    MH<R> mh = metropolis<R>(a, k_a, chain, H0, H1);

    if (a.info.proposal_momentum != nullptr) {
      // metric_vector_product with the un-normalised gradient g = dl * sL (integrators.py:136-138)
      R d = R(0);
#pragma unroll
      for (int k = 0; k < EPL; ++k) d += v[k] * dl[k];
      d = group_sum<LPC>(d) * sL;
      const R c = a2 * L * d * sL;
      R pm[EPL];
#pragma unroll
      for (int k = 0; k < EPL; ++k) pm[k] = v[k] / im[k] + c * dl[k];
      if (active) store_vec(lay, a.info.proposal_momentum, chain, pm, R(-1));
    }
    R gp[EPL];
#pragma unroll
    for (int k = 0; k < EPL; ++k) gp[k] = dl[k] * sL;  // lmc.py:226-233
    if (active) {
      store_vec(lay, a.info.proposal_position, chain, q);
      store_vec(lay, a.info.proposal_velocity, chain, v, R(-1));
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
#pragma unroll
      for (int k = 0; k < EPL; ++k) gp[k] = (g0[k] * rs0) * sL0;
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
        if (da != nullptr)
          dual_averaging_update<R>(da, mh.p_accept, (R)a.opts.da_target, (R)a.opts.da_t0, (R)a.opts.da_gamma,
                                   (R)a.opts.da_kappa);
      }
    }
    __syncwarp();  // scalars written by lane 0 are re-read by the whole group next transition
  }
}

}  // namespace gb
