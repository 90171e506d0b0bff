// Fused lmc transition (Lan et al. explicit Lagrangian integrator with volume correction).
// Reference: lmcmc/lmc.py:135-180,451-499; lmcmc/integrators.py:51-144; lmcmc/metrics.py:42-221.
//
// The reference builds dense G, jacfwd(G) (D^3), three D^3 einsums and four LU factorisations
// per step.  For the built-in metrics everything has a closed form:
//
//  * Funnel pull-back metric (examples/funnel/main.py:41-54), x = theta[:D-1], v = theta[D-1],
//    e = exp(-v), S = |x|^2, c = 1/sigma^2 -- an ARROW matrix
//        G = [[ e I, -e x / 2 ], [ -e x^T / 2, e S / 4 + c ]],   chol(G) = (J^-1)^T,
//        G^-1 = J J^T,  J = [[ e^{v/2} I, sigma x / 2 ], [ 0, sigma ]],  logdet G = -(D-1) v - 2 log sigma.
//    Omega_tilde(q, u, eps) = G + (eps/2) * (1/2)(P1 + P2 - P3) (lmcmc/metrics.py:158-179) is again an
//    arrow matrix  [[ a I, b ], [ r^T, d ]]  with
//        a = e (1 - eps u_v / 4),  r = -a x / 2,  b = -a x / 2 - eps e u_x / 4,
//        d = e S / 4 + c + (eps/2)(e x.u_x / 4 - e S u_v / 8),
//    whose determinant collapses to a^(D-1) c, so the LU log-dets of lmcmc/integrators.py:72-89 are
//        -log|det A(u, eps)| + log|det A(u', -eps)| = (D-1) (log|1 + eps u'_v / 4| - log|1 - eps u_v / 4|)
//    and the LU solve is  y_v = (w_v + x.w_x / 2) / c,  y_x = (w_x - b y_v) / a.
//  * Identity metric (metric_fn = lambda x: eye(D)): Omega_tilde = I, all volume terms vanish.
#pragma once
#include "transition.cuh"

namespace gb {

// ---- metric models: everything lmc needs from (target, metric) ---------------------------
template <typename R>
struct FunnelArrow {
  // velocity_generator lmcmc/metrics.py:75-91: u = L^-T z = J z
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                              const R (&q)[LAY::EPL], const R (&z)[LAY::EPL], R (&u)[LAY::EPL]) {
    R zl = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) zl += lay.last(k) ? z[k] : R(0);
    zl = group_sum<LAY::LPC>(zl);
    const R ev2 = fast_rsqrt(c.e);  // e^{v/2}
    const R hs = R(0.5) * tg.sigma * zl;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) u[k] = lay.last(k) ? tg.sigma * zl : fma(ev2, z[k], q[k] * hs);
  }

  // kinetic_energy lmcmc/metrics.py:93-112: -1/2 logdet G + 1/2 u^T G u
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                              const R (&q)[LAY::EPL], const R (&u)[LAY::EPL]) {
    Acc4<R, LAY::EPL> uu, xu;
    R ul = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      uu.fma(k, l ? R(0) : u[k], u[k]);
      xu.fma(k, l ? R(0) : q[k], u[k]);
      ul += l ? u[k] : R(0);
    }
    R p[3] = {uu.total(), xu.total(), ul};
    group_sum_n<LAY::LPC>(p);
    const R quad = c.e * p[0] - c.e * p[2] * p[1] + (R(0.25) * c.e * c.S + tg.inv_s2) * p[2] * p[2];
    return tg.hdm1 * c.v + log(tg.sigma) + R(0.5) * quad;
  }

  // metric_vector_product lmcmc/metrics.py:191-197
  template <class LAY>
  static __device__ __forceinline__ void Gv(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                            const R (&q)[LAY::EPL], const R (&u)[LAY::EPL], R (&p)[LAY::EPL]) {
    Acc4<R, LAY::EPL> xu;
    R ul = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      xu.fma(k, l ? R(0) : q[k], u[k]);
      ul += l ? u[k] : R(0);
    }
    R r[2] = {xu.total(), ul};
    group_sum_n<LAY::LPC>(r);
    const R pl = R(-0.5) * c.e * r[0] + (R(0.25) * c.e * c.S + tg.inv_s2) * r[1];
    const R hu = R(-0.5) * c.e * r[1];
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) p[k] = lay.last(k) ? pl : fma(c.e, u[k], q[k] * hu);
  }

  // half_step_fn lmcmc/integrators.py:61-91; returns the volume-adjustment increment
  template <class LAY>
  static __device__ __forceinline__ R half_step(const LAY& lay, const Funnel<R>& tg, const typename Funnel<R>::Ctx& c,
                                                const R (&q)[LAY::EPL], const R (&g)[LAY::EPL], R (&u)[LAY::EPL],
                                                R eps) {
    Acc4<R, LAY::EPL> xu, xg;
    R ul = R(0), gl = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      xu.fma(k, l ? R(0) : q[k], u[k]);
      xg.fma(k, l ? R(0) : q[k], g[k]);
      ul += l ? u[k] : R(0);
      gl += l ? g[k] : R(0);
    }
    R p[4] = {xu.total(), ul, xg.total(), gl};
    group_sum_n<LAY::LPC>(p);
    const R he = R(0.5) * eps, e = c.e;
    const R one_m = R(1) - R(0.25) * eps * p[1];
    const R a = e * one_m;
    // w = G u - eps/2 dphi,  dphi = -g + 1/2 grad logdet G = (-g_x, -g_v - (D-1)/2)
    const R wl = R(-0.5) * e * p[0] + (R(0.25) * e * c.S + tg.inv_s2) * p[1] + he * (p[3] + tg.hdm1);
    const R xw = e * p[0] - R(0.5) * e * c.S * p[1] + he * p[2];
    const R yl = (wl + R(0.5) * xw) * (tg.sigma * tg.sigma);
    const R ra = fast_rcp(a);
    // y_x = (w_x - b y_v) / a,  w_x = e u_x - e x u_v / 2 + (eps/2) g_x,  b = -a x / 2 - eps e u_x / 4
    const R cu = (e + R(0.25) * eps * e * yl) * ra;        // coefficient of u_x
    const R cx = R(0.5) * yl - R(0.5) * e * p[1] * ra;      // coefficient of x
    const R cg = he * ra;                                  // coefficient of g_x
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) u[k] = lay.last(k) ? yl : fma(cu, u[k], fma(cx, q[k], cg * g[k]));
    const R one_p = R(1) + R(0.25) * eps * yl;
    return R(2) * tg.hdm1 * (fast_logabs(one_p) - fast_logabs(one_m));
  }
};

template <typename R, class Target>
struct IdentityMetric {
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY&, const Target&, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&z)[LAY::EPL], R (&u)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) u[k] = z[k];
  }
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY&, const Target&, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&u)[LAY::EPL]) {
    return R(0.5) * group_sum<LAY::LPC>(dotv<R, LAY::EPL>(u, u));
  }
  template <class LAY>
  static __device__ __forceinline__ void Gv(const LAY&, const Target&, const typename Target::Ctx&,
                                            const R (&)[LAY::EPL], const R (&u)[LAY::EPL], R (&p)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) p[k] = u[k];
  }
  template <class LAY>
  static __device__ __forceinline__ R half_step(const LAY&, const Target&, const typename Target::Ctx&,
                                                const R (&)[LAY::EPL], const R (&g)[LAY::EPL], R (&u)[LAY::EPL], R eps) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) u[k] = fma(R(0.5) * eps, g[k], u[k]);
    return R(0);
  }
};

// LEAN: see lmcmonge.cuh (no Info / overrides / adaptation, legacy threefry: compiled out).
template <typename R, class Target, class Metric, int EPL, int LPC, bool EXACT, bool LEAN = false>
__global__ void __launch_bounds__(128) lmc_kernel(const TransArgs a, const Target tg) {
  using LAY = Lay<EPL, LPC, EXACT>;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long chain = tid / LPC;
  const bool active = chain < a.C;
  if (!active) chain = a.C - 1;
  LAY lay{a.D, (int)(tid % LPC)};

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

    R q[EPL], g[EPL], u[EPL];
    load_vec(lay, spos, chain, q);
    load_vec(lay, sgrad, chain, g);
    const R l0 = ((const R*)slogp)[chain];
    const R J0 = ((const R*)svol)[chain];

    U2 k_v, k_a;
    if (LEAN && LPC >= 2) {  // lean = legacy threefry: the two lanes of a pair share the key-tree blocks
      const U2 key = transition_key_shared(a, chain, t, lay.g);
      split2_shared(key, lay.g, k_v, k_a);
    } else {
      const U2 key = transition_key(a, chain, t);
      split2(LEAN ? GB200_THREEFRY_LEGACY : a.mode, key, k_v, k_a);
    }
    typename Target::Ctx ctx = tg.prepare(lay, q);
    {
      R z[EPL];
      draw_noise<R, LAY, LEAN>(a, lay, k_v, chain, z);
      Metric::draw(lay, tg, ctx, q, z, u);
      if (!LEAN && active) store_vec(lay, a.info.noise, chain, z);
    }
    if (!LEAN && active) store_vec(lay, a.info.momentum, chain, u);  // LMCInfo.velocity
    const R H0 = -l0 + Metric::kinetic(lay, tg, ctx, q, u) - J0;  // lmc_energy lmcmc/metrics.py:209-221
    R J = J0, lp = l0;

    // one_step lmcmc/integrators.py:93-142 = half-step, position + gradient refresh, half-step; written
    // as 2L half-steps so that the half-step body exists once in the instruction stream.
    // Dynamic kernels (lmcmc/lmc.py:185-252): a per-chain step count.  The warp runs to its largest count; a chain
    // that is done discards the half-step (masked commit) and moves by 0 * u, so its state stays bit-identical.
    int nh_chain = 2 * a.num_steps, nh = nh_chain;
    if (!LEAN && a.steps_per_chain != nullptr) {
      nh_chain = 2 * a.steps_per_chain[chain];
      nh = __reduce_max_sync(0xffffffffu, nh_chain);
    }
#pragma unroll 1
    for (int h = 0; h < nh; ++h) {
      const bool live = LEAN || h < nh_chain;
      const R eps_h = live ? eps : R(0);
      if (LEAN) {
        J += Metric::half_step(lay, tg, ctx, q, g, u, eps);
      } else {
        R un[EPL];
#pragma unroll
        for (int k = 0; k < EPL; ++k) un[k] = u[k];
        const R dJ = Metric::half_step(lay, tg, ctx, q, g, un, eps);
        if (live) {
#pragma unroll
          for (int k = 0; k < EPL; ++k) u[k] = un[k];
          J += dJ;
        }
      }
      if (!(h & 1)) {
#pragma unroll
        for (int k = 0; k < EPL; ++k) q[k] = fma(eps_h, u[k], q[k]);
        ctx = tg.prepare(lay, q);
        lp = tg.logp(ctx);
        tg.grad(lay, ctx, q, g);
      }
    }

    const R H1 = -lp + Metric::kinetic(lay, tg, ctx, q, u) - J;  // even in u: flip afterwards
    MH<R> mh = metropolis<R, LEAN>(a, k_a, chain, H0, H1);

    if (!LEAN && a.info.proposal_momentum != nullptr) {
      R pm[EPL];
      Metric::Gv(lay, tg, ctx, q, u, pm);
      if (active) store_vec(lay, a.info.proposal_momentum, chain, pm, R(-1));
    }
    if (!LEAN && active) {
      store_vec(lay, a.info.proposal_position, chain, q);
      store_vec(lay, a.info.proposal_velocity, chain, u, R(-1));
      store_vec(lay, a.info.proposal_logdensity_grad, chain, g);
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
      load_vec(lay, spos, chain, q);
      load_vec(lay, sgrad, chain, g);
      lp = l0;
      J = J0;
    }
    if (active) {
      store_vec(lay, a.out_pos, chain, q);
      store_vec(lay, a.out_grad, chain, g);
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
    __syncwarp();
  }
}

}  // namespace gb

namespace gb {
// Constant diagonal metric taken from the target (metric_fn returning a (D,) array: the ndim == 1
// branches of lmcmc/metrics.py:82-83,100-104,166-171,186,194): Omega_tilde = diag(G), all volume
// terms cancel, velocity = z / sqrt(G).
template <typename R, class Target>
struct TargetDiagMetric {
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&z)[LAY::EPL], R (&u)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) u[k] = z[k] * fast_rsqrt(tg.metric_diag(lay, k));
  }
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                              const R (&)[LAY::EPL], const R (&u)[LAY::EPL]) {
    R p[2] = {R(0), R(0)};
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const R G = tg.metric_diag(lay, k);
      p[0] += log(G);
      p[1] += G * u[k] * u[k];
    }
    group_sum_n<LAY::LPC>(p);
    return R(-0.5) * p[0] + R(0.5) * p[1];
  }
  template <class LAY>
  static __device__ __forceinline__ void Gv(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                            const R (&)[LAY::EPL], const R (&u)[LAY::EPL], R (&p)[LAY::EPL]) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) p[k] = tg.metric_diag(lay, k) * u[k];
  }
  template <class LAY>
  static __device__ __forceinline__ R half_step(const LAY& lay, const Target& tg, const typename Target::Ctx&,
                                                const R (&)[LAY::EPL], const R (&g)[LAY::EPL], R (&u)[LAY::EPL], R eps) {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) u[k] = fma(R(0.5) * eps, g[k] / tg.metric_diag(lay, k), u[k]);
    return R(0);
  }
};
}  // namespace gb
