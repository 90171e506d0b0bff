// Fused rmhmc transition: implicit-midpoint integrator (Brofos & Lederman) with the reference's
// plain fixed-point iteration, position-dependent metric in closed form.
// Reference: rmhmc/rmhmc.py:131-174,416-462; rmhmc/integrators.py:53-156; rmhmc/metrics.py:42-129;
//            mcmc/metrics.py:160-166 (hmc_energy).
//
// What jax.grad of the kinetic energy T(q, p) = -mvn.logpdf(p; 0, G(q)) yields (rmhmc/integrators.py:114-116):
//     dT/dp   = G^-1 p =: w
//     dT/dq_i = 1/2 tr(G^-1 d_i G) - 1/2 w^T d_i G w
// Funnel pull-back metric (examples/funnel/main.py:41-54; notation of lmc.cuh), tau := w_v:
//     w_v = sigma^2 (x.p_x / 2 + p_v),  w_x = e^{v} p_x + x tau / 2,
//     tr(G^-1 d_i G) = d_i logdet G = (0, ..., 0, -(D-1)),
//     w^T d_{x_k}G w = e tau (x_k tau / 2 - w_k),
//     w^T d_v G w    = -e |w_x|^2 + e (x.w_x) tau - e S tau^2 / 4.
#pragma once
#include "transition.cuh"

namespace gb {

template <typename R>
struct FunnelArrowH {
  // momentum_generator rmhmc/metrics.py:45-58: p = chol(G) z = (J^-1)^T z
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                              const R (&q)[LAY::EPL], const R (&z)[LAY::EPL], R (&p)[LAY::EPL]) {
    Acc4<R, LAY::EPL> xz;
    R zl = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      xz.fma(k, l ? R(0) : q[k], z[k]);
      zl += l ? z[k] : R(0);
    }
    R r[2] = {xz.total(), zl};
    group_sum_n<LAY::LPC>(r);
    const R se = sqrt(c.e);  // e^{-v/2}
    const R pl = R(-0.5) * se * r[0] + r[1] / tg.sigma;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) p[k] = lay.last(k) ? pl : se * z[k];
  }

  // w = G^-1 p (inverse_metric_vector_product rmhmc/metrics.py:120-127); returns tau = w_v
  template <class LAY>
  static __device__ __forceinline__ R Ginv(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                           const R (&q)[LAY::EPL], const R (&p)[LAY::EPL], R (&w)[LAY::EPL]) {
    Acc4<R, LAY::EPL> xp;
    R pl = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      xp.fma(k, l ? R(0) : q[k], p[k]);
      pl += l ? p[k] : R(0);
    }
    R r[2] = {xp.total(), pl};
    group_sum_n<LAY::LPC>(r);
    const R tau = tg.sigma * tg.sigma * (R(0.5) * r[0] + r[1]);
    const R ev = fast_rcp(c.e);  // e^{v}
    const R ht = R(0.5) * tau;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) w[k] = lay.last(k) ? tau : fma(ev, p[k], q[k] * ht);
    return tau;
  }

  // dT/dq given w = G^-1 p
  template <class LAY>
  static __device__ __forceinline__ void dTdq(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                              const R (&q)[LAY::EPL], const R (&w)[LAY::EPL], R tau, R (&d)[LAY::EPL]) {
    Acc4<R, LAY::EPL> ww, xw;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      ww.fma(k, l ? R(0) : w[k], w[k]);
      xw.fma(k, l ? R(0) : q[k], w[k]);
    }
    R r[2] = {ww.total(), xw.total()};
    group_sum_n<LAY::LPC>(r);
    const R e = c.e;
    const R dl = -tg.hdm1 - R(0.5) * (-e * r[0] + e * r[1] * tau - R(0.25) * e * c.S * tau * tau);
    const R c1 = R(-0.25) * e * tau * tau, c2 = R(0.5) * e * tau;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) d[k] = lay.last(k) ? dl : fma(c1, q[k], c2 * w[k]);
  }

  // kinetic_energy rmhmc/metrics.py:60-74: 1/2 p.G^-1 p + 1/2 logdet G + D/2 log(2 pi)
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                              const R (&q)[LAY::EPL], const R (&p)[LAY::EPL]) {
    R w[LAY::EPL];
    Ginv(lay, tg, c, q, p, w);
    const R pw = group_sum<LAY::LPC>(dotv<R, LAY::EPL>(p, w));
    return R(0.5) * pw - tg.hdm1 * c.v - log(tg.sigma) + R(0.91893853320467274178) * (R)lay.D();
  }
};

template <typename R, class Target>
struct IdentityMetricH {
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY&, const Target&, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&z)[LAY::EPL], R (&p)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) p[k] = z[k];
  }
  template <class LAY>
  static __device__ __forceinline__ R Ginv(const LAY&, const Target&, const typename Target::Ctx&,
                                           const R (&)[LAY::EPL], const R (&p)[LAY::EPL], R (&w)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) w[k] = p[k];
    return R(0);
  }
  template <class LAY>
  static __device__ __forceinline__ void dTdq(const LAY&, const Target&, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&)[LAY::EPL], R, R (&d)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) d[k] = R(0);
  }
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY& lay, const Target&, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&p)[LAY::EPL]) {
    return R(0.5) * group_sum<LAY::LPC>(dotv<R, LAY::EPL>(p, p)) + R(0.91893853320467274178) * (R)lay.D();
  }
};

// One evaluation of the fixed-point map (rmhmc/integrators.py:119-142):
//   (q, p) -> (qi + eps/2 dH/dp(q, p), pi - eps/2 dH/dq(q, p)),  H = T - logdensity
template <typename R, class Target, class Metric, class LAY>
__device__ __forceinline__ void midpoint_map(const LAY& lay, const Target& tg, const R (&q)[LAY::EPL],
                                             const R (&p)[LAY::EPL], const R (&qi)[LAY::EPL], const R (&pi)[LAY::EPL],
                                             R he, R (&qn)[LAY::EPL], R (&pn)[LAY::EPL]) {
  typename Target::Ctx c = tg.prepare(lay, q);
  R g[LAY::EPL], w[LAY::EPL], d[LAY::EPL];
  tg.grad(lay, c, q, g);
  const R tau = Metric::Ginv(lay, tg, c, q, p, w);
  Metric::dTdq(lay, tg, c, q, w, tau, d);
#pragma unroll
  for (int k = 0; k < LAY::EPL; ++k) {
    qn[k] = fma(he, w[k], qi[k]);
    pn[k] = fma(-he, d[k] - g[k], pi[k]);
  }
}

// LEAN: see lmcmonge.cuh (no Info / overrides / adaptation, legacy threefry: compiled out).
template <typename R, class Target, class Metric, int EPL, int LPC, bool EXACT, bool LEAN = false>
__global__ void __launch_bounds__(128) rmhmc_kernel(const TransArgs a, const Target tg) {
  using LAY = Lay<EPL, LPC, EXACT>;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long chain = tid / LPC;
  const bool active = chain < a.C;
  if (!active) chain = a.C - 1;
  LAY lay{a.D, (int)(tid % LPC)};
  const R tol = (R)a.fp_tol, div_tol = (R)a.fp_div_tol;

  const long long T = a.ks.keys ? 1 : a.ks.num_transitions;
  for (long long it = 0; it < T; ++it) {
    const long long t = a.ks.first_transition + it;
    const void* spos = it == 0 ? a.in_pos : a.out_pos;
    const void* slogp = it == 0 ? a.in_logp : a.out_logp;
    const void* sgrad = it == 0 ? a.in_grad : a.out_grad;

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
    const R he = R(0.5) * eps;

    R q[EPL], p[EPL];
    load_vec(lay, spos, chain, q);
    const R l0 = ((const R*)slogp)[chain];

    U2 key = transition_key(a, chain, t);
    U2 k_m, k_a;
    split2(LEAN ? GB200_THREEFRY_LEGACY : a.mode, key, k_m, k_a);
    typename Target::Ctx ctx = tg.prepare(lay, q);
    {
      R z[EPL];
      draw_noise<R, LAY, LEAN>(a, lay, k_m, chain, z);
      Metric::draw(lay, tg, ctx, q, z, p);
      if (!LEAN && active) store_vec(lay, a.info.noise, chain, z);
    }
    if (!LEAN && active) store_vec(lay, a.info.momentum, chain, p);  // RMHMCInfo.momentum
    const R H0 = -l0 + Metric::kinetic(lay, tg, ctx, q, p);  // hmc_energy
    int iters_total = 0;

    // Dynamic kernels (rmhmc/rmhmc.py:179-244): a per-chain step count.  The warp runs to its largest count; a chain
    // that is done steps with half step 0: the map returns its argument, the fixed point converges at once.
    int L_chain = a.num_steps, L_warp = a.num_steps;
    if (!LEAN && a.steps_per_chain != nullptr) {
      L_chain = a.steps_per_chain[chain];
      L_warp = __reduce_max_sync(0xffffffffu, L_chain);
    }
    const R he_full = he;
    for (int s = 0; s < L_warp; ++s) {  // implicit_midpoint.one_step rmhmc/integrators.py:116-154
      const R he = s < L_chain ? he_full : R(0);
      R q0[EPL], p0[EPL], qn[EPL], pn[EPL];
#pragma unroll
      for (int k = 0; k < EPL; ++k) { q0[k] = q[k]; p0[k] = p[k]; }
      // solve_fixed_point_iteration :53-89
      midpoint_map<R, Target, Metric>(lay, tg, q0, p0, q0, p0, he, q, p);
      R nrm;
      {
        R mx = R(0);
        bool nan = false;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const R dq = fabs(q[k] - q0[k]), dp = fabs(p[k] - p0[k]);
          nan = nan || isnan(dq) || isnan(dp);
          mx = fmax(mx, fmax(dq, dp));
        }
        mx = group_max<LPC>(nan ? Lim<R>::inf() : mx);
        nrm = mx;  // inf stands for "not finite" (jnp.max propagates NaN; isfinite fails either way)
      }
      int n = 0;
      for (;;) {
        const bool go = (n < a.fp_max_iters) && (nrm < Lim<R>::inf()) && (nrm < div_tol) && (nrm > tol);
        if (!__any_sync(0xffffffffu, go)) break;  // vmapped while_loop: run until every chain is done
        midpoint_map<R, Target, Metric>(lay, tg, q, p, q0, p0, he, qn, pn);
        R mx = R(0);
        bool nan = false;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
          const R dq = fabs(qn[k] - q[k]), dp = fabs(pn[k] - p[k]);
          nan = nan || isnan(dq) || isnan(dp);
          mx = fmax(mx, fmax(dq, dp));
        }
        mx = group_max<LPC>(nan ? Lim<R>::inf() : mx);
        if (go) {
#pragma unroll
          for (int k = 0; k < EPL; ++k) { q[k] = qn[k]; p[k] = pn[k]; }
          nrm = mx;
          ++n;
        }
      }
      iters_total += n;
      // explicit update from the midpoint :147-148 (initial = the midpoint itself)
      midpoint_map<R, Target, Metric>(lay, tg, q, p, q, p, he, qn, pn);
#pragma unroll
      for (int k = 0; k < EPL; ++k) { q[k] = qn[k]; p[k] = pn[k]; }
    }

    // end state: logdensity, gradient, velocity = G^-1 p; flip; energy; accept
    ctx = tg.prepare(lay, q);
    const R lp = tg.logp(ctx);
    R g[EPL];
    tg.grad(lay, ctx, q, g);
    const R H1 = -lp + Metric::kinetic(lay, tg, ctx, q, p);  // even in p
    MH<R> mh = metropolis<R, LEAN>(a, k_a, chain, H0, H1);

    if (!LEAN && a.info.proposal_velocity != nullptr) {
      R w[EPL];
      Metric::Ginv(lay, tg, ctx, q, p, w);
      if (active) store_vec(lay, a.info.proposal_velocity, chain, w, R(-1));
    }
    if (!LEAN && active) {
      store_vec(lay, a.info.proposal_position, chain, q);
      store_vec(lay, a.info.proposal_momentum, chain, p, R(-1));
      store_vec(lay, a.info.proposal_logdensity_grad, chain, g);
      if (lay.g == 0) {
        store_scalar<R>(a.info.acceptance_rate, chain, mh.p_accept);
        if (a.info.is_accepted) a.info.is_accepted[chain] = mh.accept;
        if (a.info.is_divergent) a.info.is_divergent[chain] = mh.divergent;
        store_scalar<R>(a.info.energy, chain, H1);
        store_scalar<R>(a.info.proposal_logdensity, chain, lp);
        store_scalar<R>(a.info.proposal_weight, chain, mh.weight);
        store_scalar<R>(a.info.initial_energy, chain, H0);
        store_scalar<R>(a.info.accept_uniform, chain, mh.u);
        if (a.info.fp_iters) a.info.fp_iters[chain] = iters_total;
      }
    }
    R lout = lp;
    if (!mh.accept) {
      load_vec(lay, spos, chain, q);
      load_vec(lay, sgrad, chain, g);
      lout = l0;
    }
    if (active) {
      store_vec(lay, a.out_pos, chain, q);
      store_vec(lay, a.out_grad, chain, g);
      if (a.opts.samples != nullptr)
        store_vec(lay, (R*)a.opts.samples + it * a.C * (long long)a.D, chain, q);
      if (lay.g == 0) {
        store_scalar<R>(a.out_logp, chain, lout);
        if (a.opts.sample_accept != nullptr) ((R*)a.opts.sample_accept)[it * a.C + chain] = mh.p_accept;
        if (a.opts.accept_sum != nullptr) ((R*)a.opts.accept_sum)[chain] += mh.p_accept;
        if (!LEAN && da != nullptr)
          dual_averaging_update<R>(da, mh.p_accept, (R)a.opts.da_target, (R)a.opts.da_t0, (R)a.opts.da_gamma,
                                   (R)a.opts.da_kappa);
      }
    }
    __syncwarp();
  }
}

}  // namespace gb

namespace gb {
// Constant diagonal metric from the target (ndim == 1 branches of rmhmc/metrics.py:47-48,63-65,123-124).
template <typename R, class Target>
struct TargetDiagMetricH {
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&z)[LAY::EPL], R (&p)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) p[k] = sqrt(tg.metric_diag(lay, k)) * z[k];
  }
  template <class LAY>
  static __device__ __forceinline__ R Ginv(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                           const R (&)[LAY::EPL], const R (&p)[LAY::EPL], R (&w)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) w[k] = p[k] / tg.metric_diag(lay, k);
    return R(0);
  }
  template <class LAY>
  static __device__ __forceinline__ void dTdq(const LAY&, const Target&, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&)[LAY::EPL], R, R (&d)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) d[k] = R(0);
  }
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&p)[LAY::EPL]) {
    R s[2] = {R(0), R(0)};
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const R G = tg.metric_diag(lay, k);
      s[0] += p[k] * p[k] / G;
      s[1] += log(G);
    }
    group_sum_n<LAY::LPC>(s);
    return R(0.5) * s[0] + R(0.5) * s[1] + R(0.91893853320467274178) * (R)lay.D();
  }
};
}  // namespace gb
