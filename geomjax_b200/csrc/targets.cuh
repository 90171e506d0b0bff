// In-kernel built-in targets (replace the reference's logdensity_fn / metric_fn callables).
// A D-vector of one chain is distributed over a group of LPC lanes: element j lives in lane
// j % LPC, register slot j / LPC; slots with j >= D hold zeros.
#pragma once
#include "common.cuh"

namespace gb {

template <int EPL, int LPC>
struct Lay {
  int D;  // runtime dimension, D <= EPL * LPC
  int g;  // lane within the group
  __device__ __forceinline__ int j(int k) const { return g + LPC * k; }
  __device__ __forceinline__ bool valid(int k) const { return j(k) < D; }
  __device__ __forceinline__ bool last(int k) const { return j(k) == D - 1; }
};

// Neal's funnel: examples/funnel/main.py:28-54.
//   l(theta) = N(v; 0, sigma) + sum_{i<D-1} N(x_i; 0, exp(v/2)),  v = theta[D-1]
//   grad_x = -x e^{-v};  grad_v = -v/sigma^2 - (D-1)/2 + e^{-v} S / 2,  S = sum x_i^2
//   H u    = ( e^{-v}(-u_x + x u_v),  e^{-v} x.u_x - (1/sigma^2 + e^{-v} S / 2) u_v )
template <typename R>
struct Funnel {
  R inv_s2;   // 1 / sigma^2
  R c0;       // -0.5 log(2 pi sigma^2) - (D-1)/2 log(2 pi)
  R hdm1;     // (D-1)/2
  struct Ctx { R v, e, S; };

  __host__ void setup(const gb200_target_desc& t) {
    const double s = t.params[0];
    inv_s2 = (R)(1.0 / (s * s));
    c0 = (R)(-0.5 * log(2.0 * M_PI * s * s) - 0.5 * (t.D - 1) * log(2.0 * M_PI));
    hdm1 = (R)(0.5 * (t.D - 1));
  }

  template <int EPL, int LPC>
  __device__ __forceinline__ Ctx prepare(const Lay<EPL, LPC>& lay, const R (&q)[EPL]) const {
    R p[2] = {R(0), R(0)};
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const bool l = lay.last(k);
      p[0] += l ? R(0) : q[k] * q[k];
      p[1] += l ? q[k] : R(0);
    }
    group_sum_n<LPC>(p);
    Ctx c;
    c.S = p[0];
    c.v = p[1];
    c.e = exp(-c.v);
    return c;
  }

  __device__ __forceinline__ R logp(const Ctx& c) const {
    return c0 - R(0.5) * c.v * c.v * inv_s2 - hdm1 * c.v - R(0.5) * c.e * c.S;
  }

  template <int EPL, int LPC>
  __device__ __forceinline__ void grad(const Lay<EPL, LPC>& lay, const Ctx& c, const R (&q)[EPL],
                                       R (&g)[EPL]) const {
    const R gv = -c.v * inv_s2 - hdm1 + R(0.5) * c.e * c.S;
#pragma unroll
    for (int k = 0; k < EPL; ++k) g[k] = lay.last(k) ? gv : -q[k] * c.e;
  }

  // Two Hessian-vector products with one reduction round: o1 = H u1 * s, o2 = H u2 * s.
  template <int EPL, int LPC>
  __device__ __forceinline__ void hvp2(const Lay<EPL, LPC>& lay, const Ctx& c, const R (&q)[EPL],
                                       const R (&u1)[EPL], const R (&u2)[EPL], R s,
                                       R (&o1)[EPL], R (&o2)[EPL]) const {
    R p[4] = {R(0), R(0), R(0), R(0)};  // x.u1, u1_last, x.u2, u2_last
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const bool l = lay.last(k);
      p[0] += l ? R(0) : q[k] * u1[k];
      p[1] += l ? u1[k] : R(0);
      p[2] += l ? R(0) : q[k] * u2[k];
      p[3] += l ? u2[k] : R(0);
    }
    group_sum_n<LPC>(p);
    const R es = c.e * s;
    const R hvv = -(inv_s2 + R(0.5) * c.e * c.S) * s;
    const R l1 = es * p[0] + hvv * p[1];
    const R l2 = es * p[2] + hvv * p[3];
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const bool l = lay.last(k);
      o1[k] = l ? l1 : es * (q[k] * p[1] - u1[k]);
      o2[k] = l ? l2 : es * (q[k] * p[3] - u2[k]);
    }
  }

  template <int EPL, int LPC>
  __device__ __forceinline__ void hvp(const Lay<EPL, LPC>& lay, const Ctx& c, const R (&q)[EPL],
                                      const R (&u)[EPL], R s, R (&o)[EPL]) const {
    R p[2] = {R(0), R(0)};
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
      const bool l = lay.last(k);
      p[0] += l ? R(0) : q[k] * u[k];
      p[1] += l ? u[k] : R(0);
    }
    group_sum_n<LPC>(p);
    const R es = c.e * s;
    const R hvv = -(inv_s2 + R(0.5) * c.e * c.S) * s;
    const R lv = es * p[0] + hvv * p[1];
#pragma unroll
    for (int k = 0; k < EPL; ++k) o[k] = lay.last(k) ? lv : es * (q[k] * p[1] - u[k]);
  }
};

}  // namespace gb
