// In-kernel built-in targets (replace the reference's logdensity_fn / metric_fn callables).
// A D-vector of one chain is distributed over a group of LPC lanes: element j lives in lane
// j % LPC, register slot j / LPC; slots with j >= D hold zeros.
#pragma once
#include "common.cuh"

namespace gb {

// EXACT: D == EPL * LPC is known at compile time, so every validity / "is this the last
// element" predicate folds to a constant (no ISETP/FSEL per slot).
template <int EPL_, int LPC_, bool EXACT_>
struct Lay {
  static constexpr int EPL = EPL_, LPC = LPC_;
  static constexpr bool EXACT = EXACT_;
  int D_;  // runtime dimension, D <= EPL * LPC
  int g;   // lane within the group
  __device__ __forceinline__ int D() const { return EXACT ? EPL * LPC : D_; }
  __device__ __forceinline__ int j(int k) const { return g + LPC * k; }
  __device__ __forceinline__ bool valid(int k) const { return EXACT ? true : (j(k) < D_); }
  __device__ __forceinline__ bool last(int k) const {
    if (EXACT) return (k == EPL - 1) && (LPC == 1 || g == LPC - 1);
    return j(k) == D_ - 1;
  }
};

// 4-way split accumulation: a serial chain of EPL dependent FMAs would expose 4*EPL cycles of
// latency per dot product; four partial sums cut the chain to EPL/4 + 2.
template <typename R, int EPL>
struct Acc4 {
  R a[4];
  __device__ __forceinline__ Acc4() { a[0] = a[1] = a[2] = a[3] = R(0); }
  __device__ __forceinline__ void fma(int k, R x, R y) { a[k & 3] = ::fma(x, y, a[k & 3]); }
  __device__ __forceinline__ void add(int k, R x) { a[k & 3] += x; }
  __device__ __forceinline__ R total() const { return (a[0] + a[1]) + (a[2] + a[3]); }
};

template <typename R, int EPL>
__device__ __forceinline__ R dotv(const R (&a)[EPL], const R (&b)[EPL]) {
  Acc4<R, EPL> s;
#pragma unroll
  for (int k = 0; k < EPL; ++k) s.fma(k, a[k], b[k]);
  return s.total();
}

// Neal's funnel: examples/funnel/main.py:28-54.
//   l(theta) = N(v; 0, sigma) + sum_{i<D-1} N(x_i; 0, exp(v/2)),  v = theta[D-1]
//   grad_x = -x e^{-v};  grad_v = -v/sigma^2 - (D-1)/2 + e^{-v} S / 2,  S = sum x_i^2
//   H u    = ( e^{-v}(-u_x + x u_v),  e^{-v} x.u_x - (1/sigma^2 + e^{-v} S / 2) u_v )
template <typename R>
struct Funnel {
  R inv_s2;   // 1 / sigma^2
  R sigma;
  R c0;       // -0.5 log(2 pi sigma^2) - (D-1)/2 log(2 pi)
  R hdm1;     // (D-1)/2
  struct Ctx { R v, e, S; };

  __host__ void setup(const gb200_target_desc& t) {
    const double s = t.params[0];
    inv_s2 = (R)(1.0 / (s * s));
    sigma = (R)s;
    c0 = (R)(-0.5 * log(2.0 * M_PI * s * s) - 0.5 * (t.D - 1) * log(2.0 * M_PI));
    hdm1 = (R)(0.5 * (t.D - 1));
  }

  template <class LAY>
  __device__ __forceinline__ Ctx prepare(const LAY& lay, const R (&q)[LAY::EPL]) const {
    Acc4<R, LAY::EPL> s;
    R vl = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      s.fma(k, l ? R(0) : q[k], q[k]);
      vl += l ? q[k] : R(0);
    }
    R p[2] = {s.total(), vl};
    group_sum_n<LAY::LPC>(p);
    Ctx c;
    c.S = p[0];
    c.v = p[1];
    c.e = exp(-c.v);
    return c;
  }

  __device__ __forceinline__ R logp(const Ctx& c) const {
    return c0 - R(0.5) * c.v * c.v * inv_s2 - hdm1 * c.v - R(0.5) * c.e * c.S;
  }

  // |grad|^2 from the context alone (sum_k (x_k e)^2 = e^2 S): saves the lmcmonge kernel one D-long
  // accumulation and one group reduction per integrator step; grad_scaled = grad * s in one rounding.
  static constexpr bool kGradSqnorm = true;
  __device__ __forceinline__ R grad_sqnorm(const Ctx& c) const {
    const R gv = -c.v * inv_s2 - hdm1 + R(0.5) * c.e * c.S;
    return ::fma(c.e * c.e, c.S, gv * gv);
  }
  template <class LAY>
  __device__ __forceinline__ void grad_scaled(const LAY& lay, const Ctx& c, const R (&q)[LAY::EPL], R s,
                                              R (&g)[LAY::EPL]) const {
    const R gv = (-c.v * inv_s2 - hdm1 + R(0.5) * c.e * c.S) * s;
    const R me = -c.e * s;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) g[k] = lay.last(k) ? gv : q[k] * me;
  }

  template <class LAY>
  __device__ __forceinline__ void grad(const LAY& lay, const Ctx& c, const R (&q)[LAY::EPL],
                                       R (&g)[LAY::EPL]) const {
    const R gv = -c.v * inv_s2 - hdm1 + R(0.5) * c.e * c.S;
    const R me = -c.e;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) g[k] = lay.last(k) ? gv : q[k] * me;
  }

  // Two Hessian-vector products with one reduction round: o1 = H u1 * s, o2 = H u2 * s.
  template <class LAY>
  __device__ __forceinline__ void hvp2(const LAY& lay, const Ctx& c, const R (&q)[LAY::EPL],
                                       const R (&u1)[LAY::EPL], const R (&u2)[LAY::EPL], R s,
                                       R (&o1)[LAY::EPL], R (&o2)[LAY::EPL]) const {
    Acc4<R, LAY::EPL> d1, d2;
    R l1v = R(0), l2v = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      d1.fma(k, l ? R(0) : q[k], u1[k]);
      d2.fma(k, l ? R(0) : q[k], u2[k]);
      l1v += l ? u1[k] : R(0);
      l2v += l ? u2[k] : R(0);
    }
    R p[4] = {d1.total(), l1v, d2.total(), l2v};  // x.u1, u1_last, x.u2, u2_last
    group_sum_n<LAY::LPC>(p);
    const R es = c.e * s;
    const R hvv = -(inv_s2 + R(0.5) * c.e * c.S) * s;
    const R l1 = es * p[0] + hvv * p[1];
    const R l2 = es * p[2] + hvv * p[3];
    const R a1 = es * p[1], a2 = es * p[3];
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      o1[k] = l ? l1 : ::fma(q[k], a1, -es * u1[k]);
      o2[k] = l ? l2 : ::fma(q[k], a2, -es * u2[k]);
    }
  }

  template <class LAY>
  __device__ __forceinline__ void hvp(const LAY& lay, const Ctx& c, const R (&q)[LAY::EPL],
                                      const R (&u)[LAY::EPL], R s, R (&o)[LAY::EPL]) const {
    Acc4<R, LAY::EPL> d;
    R lv = R(0);
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const bool l = lay.last(k);
      d.fma(k, l ? R(0) : q[k], u[k]);
      lv += l ? u[k] : R(0);
    }
    R p[2] = {d.total(), lv};
    group_sum_n<LAY::LPC>(p);
    const R es = c.e * s;
    const R hvv = -(inv_s2 + R(0.5) * c.e * c.S) * s;
    const R ll = es * p[0] + hvv * p[1];
    const R a1 = es * p[1];
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) o[k] = lay.last(k) ? ll : ::fma(q[k], a1, -es * u[k]);
  }
};

}  // namespace gb

namespace gb {

// Diagonal Gaussian (NEW built-in): l = -1/2 sum_j prec_j (q_j - mean_j)^2,
// constant metric diag(prec).  mean / prec are read through the read-only path on every use.
template <typename R>
struct GaussianDiag {
  const R* mean;
  const R* prec;
  struct Ctx { R quad; };
  static constexpr bool kGradSqnorm = false;

  __host__ void setup(const gb200_target_desc& t) {
    mean = (const R*)t.vec0;
    prec = (const R*)t.vec1;
  }
  template <class LAY>
  __device__ __forceinline__ R mu(const LAY& lay, int k) const { return lay.valid(k) ? __ldg(mean + lay.j(k)) : R(0); }
  template <class LAY>
  __device__ __forceinline__ R metric_diag(const LAY& lay, int k) const { return lay.valid(k) ? __ldg(prec + lay.j(k)) : R(1); }

  template <class LAY>
  __device__ __forceinline__ Ctx prepare(const LAY& lay, const R (&q)[LAY::EPL]) const {
    Acc4<R, LAY::EPL> s;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      const R d = lay.valid(k) ? q[k] - mu(lay, k) : R(0);
      s.fma(k, d * metric_diag(lay, k), d);
    }
    Ctx c;
    c.quad = group_sum<LAY::LPC>(s.total());
    return c;
  }
  __device__ __forceinline__ R logp(const Ctx& c) const { return R(-0.5) * c.quad; }
  template <class LAY>
  __device__ __forceinline__ void grad(const LAY& lay, const Ctx&, const R (&q)[LAY::EPL], R (&g)[LAY::EPL]) const {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) g[k] = lay.valid(k) ? -metric_diag(lay, k) * (q[k] - mu(lay, k)) : R(0);
  }
  template <class LAY>
  __device__ __forceinline__ void hvp(const LAY& lay, const Ctx&, const R (&)[LAY::EPL], const R (&u)[LAY::EPL], R s,
                                      R (&o)[LAY::EPL]) const {
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) o[k] = lay.valid(k) ? -metric_diag(lay, k) * u[k] * s : R(0);
  }
  template <class LAY>
  __device__ __forceinline__ void hvp2(const LAY& lay, const Ctx& c, const R (&q)[LAY::EPL], const R (&u1)[LAY::EPL],
                                       const R (&u2)[LAY::EPL], R s, R (&o1)[LAY::EPL], R (&o2)[LAY::EPL]) const {
    hvp(lay, c, q, u1, s, o1);
    hvp(lay, c, q, u2, s, o2);
  }
};

// Banana (NEW built-in; Haario's twisted Gaussian), D = 2:
//   l = -x1^2 / (2 s1) - (x2 - b (x1^2 - s1))^2 / 2,  identity metric.
template <typename R>
struct Banana {
  R s1, inv_s1, b;
  struct Ctx { R x1, x2, r; };
  static constexpr bool kGradSqnorm = false;

  __host__ void setup(const gb200_target_desc& t) {
    s1 = (R)t.params[0];
    inv_s1 = (R)(1.0 / t.params[0]);
    b = (R)t.params[1];
  }
  template <class LAY>
  __device__ __forceinline__ void pair(const LAY& lay, const R (&v)[LAY::EPL], R& a0, R& a1) const {
    R p[2] = {R(0), R(0)};
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) {
      p[0] += (lay.j(k) == 0) ? v[k] : R(0);
      p[1] += (lay.j(k) == 1) ? v[k] : R(0);
    }
    group_sum_n<LAY::LPC>(p);
    a0 = p[0];
    a1 = p[1];
  }
  template <class LAY>
  __device__ __forceinline__ Ctx prepare(const LAY& lay, const R (&q)[LAY::EPL]) const {
    Ctx c;
    pair(lay, q, c.x1, c.x2);
    c.r = c.x2 - b * (c.x1 * c.x1 - s1);
    return c;
  }
  __device__ __forceinline__ R logp(const Ctx& c) const { return -c.x1 * c.x1 * (R(0.5) * inv_s1) - R(0.5) * c.r * c.r; }
  template <class LAY>
  __device__ __forceinline__ void grad(const LAY& lay, const Ctx& c, const R (&)[LAY::EPL], R (&g)[LAY::EPL]) const {
    const R g0 = -c.x1 * inv_s1 + R(2) * b * c.x1 * c.r;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) g[k] = (lay.j(k) == 0) ? g0 : ((lay.j(k) == 1) ? -c.r : R(0));
  }
  template <class LAY>
  __device__ __forceinline__ void hvp(const LAY& lay, const Ctx& c, const R (&)[LAY::EPL], const R (&u)[LAY::EPL], R s,
                                      R (&o)[LAY::EPL]) const {
    R u1, u2;
    pair(lay, u, u1, u2);
    const R h00 = -inv_s1 + R(2) * b * c.r - R(4) * b * b * c.x1 * c.x1, h01 = R(2) * b * c.x1;
    const R o0 = (h00 * u1 + h01 * u2) * s, o1 = (h01 * u1 - u2) * s;
#pragma unroll
    for (int k = 0; k < LAY::EPL; ++k) o[k] = (lay.j(k) == 0) ? o0 : ((lay.j(k) == 1) ? o1 : R(0));
  }
  template <class LAY>
  __device__ __forceinline__ void hvp2(const LAY& lay, const Ctx& c, const R (&q)[LAY::EPL], const R (&u1)[LAY::EPL],
                                       const R (&u2)[LAY::EPL], R s, R (&o1)[LAY::EPL], R (&o2)[LAY::EPL]) const {
    hvp(lay, c, q, u1, s, o1);
    hvp(lay, c, q, u2, s, o2);
  }
};

}  // namespace gb
