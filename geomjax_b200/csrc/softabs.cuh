// SoftAbs metric (Betancourt 2013) for Neal's funnel in D = 2: the metric
// BASELINE.json's configs[0] names for rmhmc.  NEW (not in the reference): the contract is the
// reference's metric_fn(theta) -> (D, D) SPD consumed by rmhmc/metrics.py:42-129; specification =
// oracle/targets.py::softabs_metric.
//   H = -hessian(logp) = Q diag(lam) Q^T,  G = Q diag(f(lam)) Q^T,  f(lam) = lam coth(alpha lam)
//   d_k G = Q (J o (Q^T d_k H Q)) Q^T,  J_ij = (f_i - f_j)/(lam_i - lam_j),  J_ii = f'(lam_i)
// The Hessian entries are formed in float32 exactly as the oracle forms them, the 2x2 eigen-system
// and f run in float64 and G, dG are rounded to float32 once (the oracle does the same through
// numpy.linalg.eigh); Cholesky / solve / traces are float32 closed forms.  One thread per chain.
#pragma once
#include "rmhmc.cuh"

namespace gb {

template <typename R>
struct FunnelSA : Funnel<R> {
  double alpha;
  struct Ctx : Funnel<R>::Ctx {
    R G00, G01, G11;   // metric
    R A00, A01, A11;   // d G / d x
    R B00, B01, B11;   // d G / d v
  };

  __host__ void setup(const gb200_target_desc& t) {
    Funnel<R>::setup(t);
    alpha = t.params[7];
  }

  static __device__ __forceinline__ void f_and_fp(double alpha, double lam, double& f, double& fp) {
    const double x = alpha * lam;
    if (fabs(x) < 1e-4) {
      f = (1.0 + x * x / 3.0) / alpha;
      fp = (2.0 / 3.0) * x;
    } else if (fabs(x) > 30.0) {
      f = fabs(lam);
      fp = x > 0 ? 1.0 : -1.0;
    } else {
      const double coth = 1.0 / tanh(x), sh = sinh(x);
      f = lam * coth;
      fp = coth - x / (sh * sh);
    }
  }

  template <class LAY>
  __device__ __noinline__ Ctx prepare(const LAY& lay, const R (&q)[LAY::EPL]) const {
    static_assert(LAY::EPL == 2 && LAY::LPC == 1, "SoftAbs is built for D = 2, one thread per chain");
    Ctx c;
    const R x = q[0], v = q[1];
    c.v = v;
    c.S = x * x;
    c.e = exp(-v);
    // -hessian(logp) and its derivatives, float32 entries as the oracle forms them
    const float e = (float)c.e, xf = (float)x;
    const float ex = __fmul_rn(xf, e);
    const float hx2 = __fmul_rn(__fmul_rn(0.5f, e), __fmul_rn(xf, xf));
    const double a = (double)e, b = -(double)ex, d = (double)__fadd_rn((float)this->inv_s2, hx2);
    // d_x H = [[0, -e], [-e, x e]];  d_v H = [[-e, x e], [x e, -e x^2 / 2]]
    const double ax00 = 0.0, ax01 = -(double)e, ax11 = (double)ex;
    const double av00 = -(double)e, av01 = (double)ex, av11 = -(double)hx2;
    // symmetric 2x2 eigen-system
    const double m = 0.5 * (a + d), dd = 0.5 * (a - d);
    const double r = hypot(dd, b);
    const double l1 = m + r, l2 = m - r;
    double cx = 1.0, sx = 0.0;
    if (r > 0.0) {
      double vx, vy;
      if (dd >= 0.0) { vx = dd + r; vy = b; } else { vx = b; vy = r - dd; }
      const double n = hypot(vx, vy);
      cx = vx / n;
      sx = vy / n;
    }
    // v1 = (cx, sx), v2 = (-sx, cx)
    double f1, f2, g1, g2;
    f_and_fp(alpha, l1, f1, g1);
    f_and_fp(alpha, l2, f2, g2);
    double j12;
    if (fabs(l1 - l2) <= 1e-9 * fmax(fabs(l1), fabs(l2)) + 1e-300) {
      double fm;
      f_and_fp(alpha, 0.5 * (l1 + l2), fm, j12);
    } else {
      j12 = (f1 - f2) / (l1 - l2);
    }
    const double p11 = cx * cx, p12 = cx * sx, p22 = sx * sx;  // v1 v1^T; v2 v2^T = [[p22, -p12], [-p12, p11]]
    c.G00 = (R)(f1 * p11 + f2 * p22);
    c.G01 = (R)((f1 - f2) * p12);
    c.G11 = (R)(f1 * p22 + f2 * p11);
    // v1 v2^T + v2 v1^T = [[-2 p12, p11 - p22], [p11 - p22, 2 p12]]
    const double s00 = -2.0 * p12, s01 = p11 - p22, s11 = 2.0 * p12;
#define GB_SA_D(o00, o01, o11, h00, h01, h11)                                               \
  {                                                                                         \
    const double m11 = h00 * p11 + 2.0 * h01 * p12 + h11 * p22;   /* v1^T dH v1 */           \
    const double m22 = h00 * p22 - 2.0 * h01 * p12 + h11 * p11;   /* v2^T dH v2 */           \
    const double m12 = (h11 - h00) * p12 + h01 * (p11 - p22);     /* v1^T dH v2 */           \
    const double c1 = g1 * m11, c2 = g2 * m22, c3 = j12 * m12;                              \
    o00 = (R)(c1 * p11 + c2 * p22 + c3 * s00);                                              \
    o01 = (R)((c1 - c2) * p12 + c3 * s01);                                                  \
    o11 = (R)(c1 * p22 + c2 * p11 + c3 * s11);                                              \
  }
    GB_SA_D(c.A00, c.A01, c.A11, ax00, ax01, ax11)
    GB_SA_D(c.B00, c.B01, c.B11, av00, av01, av11)
#undef GB_SA_D
    return c;
  }
};

// Dense 2x2 metric carried by the target context (rmhmc/metrics.py:42-129, ndim == 2 branches).
template <typename R>
struct SoftAbs2H {
  using Tg = FunnelSA<R>;
  using Ctx = typename Tg::Ctx;
  template <class LAY>
  static __device__ __forceinline__ void draw(const LAY&, const Tg&, const Ctx& c, const R (&)[2], const R (&z)[2], R (&p)[2]) {
    const R l00 = sqrt(c.G00), l10 = c.G01 / l00, l11 = sqrt(c.G11 - l10 * l10);  // momentum_generator :45-58
    p[0] = l00 * z[0];
    p[1] = fma(l10, z[0], l11 * z[1]);
  }
  template <class LAY>
  static __device__ __forceinline__ R Ginv(const LAY&, const Tg&, const Ctx& c, const R (&)[2], const R (&p)[2], R (&w)[2]) {
    const R det = c.G00 * c.G11 - c.G01 * c.G01;  // inverse_metric_vector_product :120-127
    w[0] = (c.G11 * p[0] - c.G01 * p[1]) / det;
    w[1] = (c.G00 * p[1] - c.G01 * p[0]) / det;
    return R(0);
  }
  template <class LAY>
  static __device__ __forceinline__ void dTdq(const LAY&, const Tg&, const Ctx& c, const R (&)[2], const R (&w)[2], R, R (&d)[2]) {
    const R det = c.G00 * c.G11 - c.G01 * c.G01;
    const R trA = (c.G11 * c.A00 - R(2) * c.G01 * c.A01 + c.G00 * c.A11) / det;
    const R trB = (c.G11 * c.B00 - R(2) * c.G01 * c.B01 + c.G00 * c.B11) / det;
    const R qA = c.A00 * w[0] * w[0] + R(2) * c.A01 * w[0] * w[1] + c.A11 * w[1] * w[1];
    const R qB = c.B00 * w[0] * w[0] + R(2) * c.B01 * w[0] * w[1] + c.B11 * w[1] * w[1];
    d[0] = R(0.5) * trA - R(0.5) * qA;
    d[1] = R(0.5) * trB - R(0.5) * qB;
  }
  template <class LAY>
  static __device__ __forceinline__ R kinetic(const LAY&, const Tg&, const Ctx& c, const R (&)[2], const R (&p)[2]) {
    const R l00 = sqrt(c.G00), l10 = c.G01 / l00, l11 = sqrt(c.G11 - l10 * l10);  // kinetic_energy :60-74
    const R y0 = p[0] / l00, y1 = (p[1] - l10 * y0) / l11;
    return R(0.5) * (y0 * y0 + y1 * y1) + log(l00) + log(l11) + R(2.0 * 0.91893853320467274178);
  }
};

}  // namespace gb
