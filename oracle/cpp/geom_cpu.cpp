// CPU restatement (C++ / OpenMP over chains) of geomjax's static transition kernels -- ORACLE / CPU BASELINE ONLY.
// TEST INFRASTRUCTURE: used by tests/ and bench.py's cpu_baseline / --impl reference legs; never by geomjax_b200/.
//
// Reference semantics restated (file:line under /root/reference/geomjax):
//   lmcmonge  lmcmonge/lmc.py:151-235,512-565; lmcmonge/integrators.py:52-230; lmcmonge/metrics.py:44-284
//   lmc       lmcmc/lmc.py:135-180,451-499;   lmcmc/integrators.py:51-144;  lmcmc/metrics.py:42-221
//   rmhmc     rmhmc/rmhmc.py:131-174,416-462; rmhmc/integrators.py:53-156;  rmhmc/metrics.py:42-129
//   shared    mcmc/trajectory.py:121-139; mcmc/proposal.py:43-121,168-185; mcmc/metrics.py:160-166; util.py:58-83
//   targets   examples/funnel/main.py:28-54 (Neal's funnel + pull-back metric); logistic regression + Fisher
//             metric is the NEW target of SURVEY Appendix B.1.
// As BASELINE.md section 4(a) specifies, autodiff is replaced by analytic derivatives and the dense D^3 algebra of
// the funnel's pull-back metric by its closed (arrow-matrix) forms; the Monge draw keeps the reference's dense
// Cholesky and logistic regression keeps dense per-chain Cholesky / inverse.  float32 throughout (JAX without
// x64).  Checked against the NumPy oracle (oracle/samplers.py) in tests/test_oracle_cpp.py.
#include "geom_cpu.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace ocpu {

static const float HALF_LOG_2PI = 0.91893853320467274178f;

struct Out {  // per-chain outputs of one transition
  float accept_rate, energy, initial_energy, u;
  bool accepted;
  int fp_iters;
};

struct Scratch {
  std::vector<float> buf;
  float* take(size_t n, size_t& off) {
    float* p = buf.data() + off;
    off += n;
    return p;
  }
};

static inline float dotf(const float* a, const float* b, int n) {
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// mcmc/proposal.py:87-121 + :168-185
static void metropolis(const ocpu_problem& P, Key k_a, float H0, float H1, Out* o) {
  float delta = H0 - H1;
  if (std::isnan(delta)) delta = -INFINITY;
  o->accept_rate = std::fmin(std::exp(delta), 1.0f);
  o->u = uniform01(k_a);
  o->accepted = o->u < o->accept_rate;
  o->energy = H1;
  o->initial_energy = H0;
}

// ------------------------------------------------------------------------------------------------- funnel
struct Funnel {
  int D;
  float sigma;
  // examples/funnel/main.py:34-39 with jax.scipy.stats.norm.logpdf
  float logp(const float* q) const {
    const float v = q[D - 1], two_pi = 6.28318530717958647692f, s2 = sigma * sigma;
    const float top = -(std::log(two_pi * s2) + v * v / s2) / 2.f;
    const float scale = std::exp(0.5f * v), sc2 = scale * scale, lg = std::log(two_pi * sc2);
    float s = 0.f;
    for (int k = 0; k < D - 1; ++k) s += -(lg + q[k] * q[k] / sc2) / 2.f;
    return top + s;
  }
  void grad(const float* q, float* g) const {
    const float v = q[D - 1], e = std::exp(-v);
    float S = 0.f;
    for (int k = 0; k < D - 1; ++k) {
      g[k] = -q[k] * e;
      S += q[k] * q[k];
    }
    g[D - 1] = -v / (sigma * sigma) - 0.5f * (float)(D - 1) + 0.5f * e * S;
  }
  void hvp(const float* q, const float* u, float* out) const {
    const float v = q[D - 1], e = std::exp(-v), uv = u[D - 1];
    float S = 0.f, xu = 0.f;
    for (int k = 0; k < D - 1; ++k) {
      out[k] = e * (-u[k] + q[k] * uv);
      S += q[k] * q[k];
      xu += q[k] * u[k];
    }
    const float hvv = -1.0f / (sigma * sigma) - 0.5f * e * S;
    out[D - 1] = e * xu + hvv * uv;
  }
};

// ------------------------------------------------------------------------------------------------- lmcmonge
struct MongeState {
  float *q, *v, *dl, *dl_ig, *Hdl_ig, *ig_Hdl_ig, *Hv;
  float l, L, J;
};

// lmcmonge/integrators.py:158-230 (`omega` as written: no alpha2 on the Christoffel correction)
static void monge_half_step(const ocpu_problem& P, int D, MongeState& s, float eps, float* vt) {
  const float a2 = (float)P.alpha2, half = 0.5f;
  const float det1 = 1.f + half * eps * a2 * dotf(s.Hv, s.dl_ig, D);
  s.J = s.J - std::log(std::fabs(det1));
  if (P.half_step == OCPU_OMEGA || P.half_step == OCPU_OMEGA_FIXED) {
    const float sL = std::sqrt(s.L);
    float dphi_dlig = 0.f;  // dot(dphi, dl_ig), dphi = -dl + a2 Hdl_ig / sL
    for (int i = 0; i < D; ++i) dphi_dlig += (-s.dl[i] + a2 * s.Hdl_ig[i] / sL) * s.dl_ig[i];
    const float cc = a2 * dphi_dlig, hs = half * eps * sL;
    for (int i = 0; i < D; ++i) {
      const float dphi_ig = -s.dl_ig[i] + a2 * s.ig_Hdl_ig[i] / sL;
      vt[i] = s.v[i] - hs * (dphi_ig - cc * s.dl_ig[i]);
    }
    float c = half * eps * dotf(vt, s.Hv, D) / det1;
    if (P.half_step == OCPU_OMEGA_FIXED) c = a2 * c;
    for (int i = 0; i < D; ++i) s.v[i] = vt[i] - c * s.dl_ig[i];
  } else {
    const float c1 = a2 * s.L * dotf(s.dl, s.v, D) + half * eps * std::sqrt(s.L);
    for (int i = 0; i < D; ++i) vt[i] = s.v[i] + c1 * s.dl_ig[i] - half * a2 * eps * s.ig_Hdl_ig[i];
    const float num = a2 * (dotf(s.dl, vt, D) + half * eps * dotf(s.Hv, vt, D));
    const float c = num / det1;
    for (int i = 0; i < D; ++i) s.v[i] = vt[i] - c * s.dl_ig[i];
  }
  s.J = s.J + std::log(std::fabs(1.f - half * eps * a2 * dotf(s.Hdl_ig, s.v, D)));
}

static float monge_kinetic(const ocpu_problem& P, int D, const float* mass, const float* v, const float* dl, float L) {
  float lm = 0.f, vv = 0.f;
  for (int i = 0; i < D; ++i) {
    lm += std::log(mass[i]);
    const float vg = std::sqrt(mass[i]) * v[i];
    vv += vg * vg;
  }
  const float vd = dotf(v, dl, D);
  return -0.5f * (std::log(L) + lm) + 0.5f * vv + 0.5f * L * (float)P.alpha2 * (vd * vd);
}

static void lmcmonge_chain(const ocpu_problem& P, Key key, float* q0, float* l0, float* g0, float* J0, Scratch& S,
                           Out* o, float* draw_out, float* prop_out) {
  const int D = P.D;
  const Funnel tg{D, (float)P.sigma};
  const float a2 = (float)P.alpha2, eps = (float)P.step_size;
  size_t off = 0;
  float* inv_mass = S.take(D, off);
  float* mass = S.take(D, off);
  for (int i = 0; i < D; ++i) {
    inv_mass[i] = P.inv_mass ? P.inv_mass[i] : 1.f;
    mass[i] = 1.f / inv_mass[i];
  }
  MongeState s;
  s.q = S.take(D, off); s.v = S.take(D, off); s.dl = S.take(D, off); s.dl_ig = S.take(D, off);
  s.Hdl_ig = S.take(D, off); s.ig_Hdl_ig = S.take(D, off); s.Hv = S.take(D, off);
  float* z = S.take(D, off);
  float* tmp = S.take(D, off);
  float* g = S.take(D, off);
  float* dl0 = S.take(D, off);
  float* A = S.take((size_t)D * D, off);
  const Key k_v = split_index(key, 2, 0), k_a = split_index(key, 2, 1);
  // lmcmonge/lmc.py:177-208
  float gg = 0.f;
  for (int i = 0; i < D; ++i) gg += inv_mass[i] * g0[i] * g0[i];
  const float L0 = 1.f + a2 * gg, sL0 = std::sqrt(L0);
  for (int i = 0; i < D; ++i) {
    s.q[i] = q0[i];
    s.dl[i] = g0[i] / sL0;
    dl0[i] = s.dl[i];
    s.dl_ig[i] = inv_mass[i] * s.dl[i];
  }
  tg.hvp(q0, s.dl_ig, s.Hdl_ig);
  for (int i = 0; i < D; ++i) {
    s.Hdl_ig[i] /= sL0;
    s.ig_Hdl_ig[i] = inv_mass[i] * s.Hdl_ig[i];
  }
  normal(k_v, D, z);
  // velocity_generator lmcmonge/metrics.py:155-166: dense Cholesky of G^-1 = diag(inv_mass) - a2 dl_ig dl_ig^T
  for (int i = 0; i < D; ++i)
    for (int j = 0; j <= i; ++j) A[i * D + j] = (i == j ? inv_mass[i] : 0.f) - a2 * s.dl_ig[i] * s.dl_ig[j];
  for (int j = 0; j < D; ++j) {
    float d = A[j * D + j];
    for (int k = 0; k < j; ++k) d -= A[j * D + k] * A[j * D + k];
    d = std::sqrt(d);
    A[j * D + j] = d;
    for (int i = j + 1; i < D; ++i) {
      float t = A[i * D + j];
      for (int k = 0; k < j; ++k) t -= A[i * D + k] * A[j * D + k];
      A[i * D + j] = t / d;
    }
  }
  for (int i = 0; i < D; ++i) {
    float t = 0.f;
    for (int j = 0; j <= i; ++j) t += A[i * D + j] * z[j];
    s.v[i] = t;
    if (draw_out) draw_out[i] = t;
  }
  tg.hvp(q0, s.v, s.Hv);
  for (int i = 0; i < D; ++i) s.Hv[i] /= sL0;
  s.l = *l0; s.L = L0; s.J = *J0;
  const float H0 = -s.l + monge_kinetic(P, D, mass, s.v, s.dl, s.L) - s.J;
  for (int step = 0; step < P.L; ++step) {  // lmcmonge/integrators.py:63-153
    monge_half_step(P, D, s, eps, tmp);
    for (int i = 0; i < D; ++i) s.q[i] += eps * s.v[i];
    s.l = tg.logp(s.q);
    tg.grad(s.q, g);
    float g2 = 0.f;
    for (int i = 0; i < D; ++i) g2 += inv_mass[i] * g[i] * g[i];
    s.L = 1.f + a2 * g2;
    const float sL = std::sqrt(s.L);
    for (int i = 0; i < D; ++i) s.dl[i] = g[i] / sL;
    tg.hvp(s.q, s.v, s.Hv);
    for (int i = 0; i < D; ++i) {
      s.Hv[i] /= sL;
      s.dl_ig[i] = inv_mass[i] * s.dl[i];
    }
    tg.hvp(s.q, s.dl_ig, s.Hdl_ig);
    for (int i = 0; i < D; ++i) {
      s.Hdl_ig[i] /= sL;
      s.ig_Hdl_ig[i] = inv_mass[i] * s.Hdl_ig[i];
    }
    monge_half_step(P, D, s, eps, tmp);
    tg.hvp(s.q, s.v, s.Hv);
    for (int i = 0; i < D; ++i) s.Hv[i] /= sL;
  }
  for (int i = 0; i < D; ++i) s.v[i] = -s.v[i];  // flip_velocity lmcmonge/lmc.py:539-565
  const float H1 = -s.l + monge_kinetic(P, D, mass, s.v, s.dl, s.L) - s.J;
  metropolis(P, k_a, H0, H1, o);
  if (prop_out) std::memcpy(prop_out, s.q, sizeof(float) * D);
  // lmcmonge/lmc.py:226-233: the gradient is re-scaled from the normalised one of the sampled state
  if (o->accepted) {
    const float sL = std::sqrt(s.L);
    for (int i = 0; i < D; ++i) {
      q0[i] = s.q[i];
      g0[i] = s.dl[i] * sL;
    }
    *l0 = s.l;
    *J0 = s.J;
  } else {
    for (int i = 0; i < D; ++i) g0[i] = dl0[i] * sL0;
  }
}

// ------------------------------------------------------------------------------------------------- lmc (funnel)
// The funnel's pull-back metric is an arrow matrix; Omega_tilde, its LU solve and log-determinants have the closed
// forms derived in geomjax_b200/csrc/lmc.cuh (header) -- the same algebra the CUDA path runs.
struct Arrow {
  int D;
  float sigma, c, hdm1;  // c = 1/sigma^2, hdm1 = (D-1)/2
  float e, S, v;         // context at the current position
  void ctx(const float* q) {
    v = q[D - 1];
    e = std::exp(-v);
    S = 0.f;
    for (int k = 0; k < D - 1; ++k) S += q[k] * q[k];
  }
  float kinetic(const float* q, const float* u) const {  // lmcmc/metrics.py:93-112
    float uu = 0.f, xu = 0.f;
    for (int k = 0; k < D - 1; ++k) {
      uu += u[k] * u[k];
      xu += q[k] * u[k];
    }
    const float ul = u[D - 1];
    const float quad = e * uu - e * ul * xu + (0.25f * e * S + c) * ul * ul;
    return hdm1 * v + std::log(sigma) + 0.5f * quad;
  }
  float half_step(const float* q, const float* g, float* u, float eps) const {  // lmcmc/integrators.py:61-91
    float xu = 0.f, xg = 0.f;
    for (int k = 0; k < D - 1; ++k) {
      xu += q[k] * u[k];
      xg += q[k] * g[k];
    }
    const float ul = u[D - 1], gl = g[D - 1], he = 0.5f * eps;
    const float one_m = 1.f - 0.25f * eps * ul, a = e * one_m;
    const float wl = -0.5f * e * xu + (0.25f * e * S + c) * ul + he * (gl + hdm1);
    const float xw = e * xu - 0.5f * e * S * ul + he * xg;
    const float yl = (wl + 0.5f * xw) * (sigma * sigma);
    const float cu = (e + 0.25f * eps * e * yl) / a, cx = 0.5f * yl - 0.5f * e * ul / a, cg = he / a;
    for (int k = 0; k < D - 1; ++k) u[k] = cu * u[k] + cx * q[k] + cg * g[k];
    u[D - 1] = yl;
    const float one_p = 1.f + 0.25f * eps * yl;
    return 2.f * hdm1 * (std::log(std::fabs(one_p)) - std::log(std::fabs(one_m)));
  }
};

static void lmc_chain(const ocpu_problem& P, Key key, float* q0, float* l0, float* g0, float* J0, Scratch& S, Out* o,
                      float* draw_out, float* prop_out) {
  const int D = P.D;
  const Funnel tg{D, (float)P.sigma};
  Arrow ar{D, (float)P.sigma, 1.f / (float)(P.sigma * P.sigma), 0.5f * (float)(D - 1), 0.f, 0.f, 0.f};
  const float eps = (float)P.step_size;
  size_t off = 0;
  float* q = S.take(D, off);
  float* u = S.take(D, off);
  float* g = S.take(D, off);
  float* z = S.take(D, off);
  const Key k_v = split_index(key, 2, 0), k_a = split_index(key, 2, 1);
  normal(k_v, D, z);
  ar.ctx(q0);
  // velocity_generator lmcmc/metrics.py:75-91: u = chol(G)^-T z = J z, J = [[e^{v/2} I, sigma x / 2], [0, sigma]]
  const float ev2 = std::exp(0.5f * ar.v), zl = z[D - 1];
  for (int k = 0; k < D - 1; ++k) u[k] = ev2 * z[k] + q0[k] * (0.5f * ar.sigma * zl);
  u[D - 1] = ar.sigma * zl;
  if (draw_out) std::memcpy(draw_out, u, sizeof(float) * D);
  std::memcpy(q, q0, sizeof(float) * D);
  std::memcpy(g, g0, sizeof(float) * D);
  float l = *l0, J = *J0;
  const float H0 = -l + ar.kinetic(q, u) - J;  // lmc_energy lmcmc/metrics.py:209-221
  for (int step = 0; step < P.L; ++step) {       // lmcmc/integrators.py:93-142
    J += ar.half_step(q, g, u, eps);
    for (int i = 0; i < D; ++i) q[i] += eps * u[i];
    l = tg.logp(q);
    tg.grad(q, g);
    ar.ctx(q);
    J += ar.half_step(q, g, u, eps);
  }
  for (int i = 0; i < D; ++i) u[i] = -u[i];  // flip_velocity lmcmc/lmc.py:478-499
  const float H1 = -l + ar.kinetic(q, u) - J;
  metropolis(P, k_a, H0, H1, o);
  if (prop_out) std::memcpy(prop_out, q, sizeof(float) * D);
  if (o->accepted) {
    std::memcpy(q0, q, sizeof(float) * D);
    std::memcpy(g0, g, sizeof(float) * D);
    *l0 = l;
    *J0 = J;
  }
}

// ------------------------------------------------------------------------------------------------- rmhmc (logreg)
struct LogReg {
  int N, D;
  float alpha;
  const float *X, *y;
  float *s, *G, *Gi, *h;  // scratch: s [N], G (Cholesky factor, lower) [D, D], Gi = G^-1 [D, D], h [N]

  float logp(const float* q) const {
    float acc = 0.f, qq = 0.f;
    for (int n = 0; n < N; ++n) {
      const float eta = dotf(X + (size_t)n * D, q, D);
      acc += y[n] * eta - (std::fmax(eta, 0.f) + std::log1p(std::exp(-std::fabs(eta))));  // logaddexp(0, eta)
    }
    for (int i = 0; i < D; ++i) qq += q[i] * q[i];
    return acc - 0.5f * alpha * qq;
  }
  // s = sigmoid(X q); grad = X^T (y - s) - alpha q
  void sig_grad(const float* q, float* g) {
    for (int i = 0; i < D; ++i) g[i] = 0.f;
    for (int n = 0; n < N; ++n) {
      const float* x = X + (size_t)n * D;
      const float sg = 1.f / (1.f + std::exp(-dotf(x, q, D)));
      s[n] = sg;
      const float r = y[n] - sg;
      for (int i = 0; i < D; ++i) g[i] += r * x[i];
    }
    for (int i = 0; i < D; ++i) g[i] -= alpha * q[i];
  }
  // G = X^T diag(s(1-s)) X + alpha I -> lower Cholesky factor in G (needs s); returns sum(log diag)
  float factor() {
    for (int i = 0; i < D * D; ++i) G[i] = 0.f;
    for (int n = 0; n < N; ++n) {
      const float* x = X + (size_t)n * D;
      const float w = s[n] * (1.f - s[n]);
      for (int i = 0; i < D; ++i) {
        const float a = w * x[i];
        float* Gr = G + (size_t)i * D;
        for (int j = 0; j <= i; ++j) Gr[j] += a * x[j];
      }
    }
    for (int i = 0; i < D; ++i) G[i * D + i] += alpha;
    float ld = 0.f;
    for (int j = 0; j < D; ++j) {
      float d = G[j * D + j];
      for (int k = 0; k < j; ++k) d -= G[j * D + k] * G[j * D + k];
      d = std::sqrt(d);
      G[j * D + j] = d;
      ld += std::log(d);
      for (int i = j + 1; i < D; ++i) {
        float t = G[i * D + j];
        for (int k = 0; k < j; ++k) t -= G[i * D + k] * G[j * D + k];
        G[i * D + j] = t / d;
      }
    }
    return ld;
  }
  void lower_solve(const float* b, float* yv) const {  // L yv = b
    for (int i = 0; i < D; ++i) {
      float t = b[i];
      for (int k = 0; k < i; ++k) t -= G[i * D + k] * yv[k];
      yv[i] = t / G[i * D + i];
    }
  }
  void upper_solve(float* yv) const {  // L^T x = yv, in place
    for (int i = D - 1; i >= 0; --i) {
      float t = yv[i];
      for (int k = i + 1; k < D; ++k) t -= G[k * D + i] * yv[k];
      yv[i] = t / G[i * D + i];
    }
  }
  // dT/dq_i = 1/2 sum_n w'_n x_ni (x_n^T G^-1 x_n - (x_n . v)^2), v = G^-1 p (rmhmc/integrators.py:114-116 with
  // d_i G = X^T diag(w' x_.i) X); needs factor()
  void kinetic_grad(const float* p, float* v, float* dT, float* tmp) {
    lower_solve(p, v);
    upper_solve(v);
    // G^-1 = L^-T L^-1: columns by two triangular solves
    for (int c = 0; c < D; ++c) {
      for (int i = 0; i < D; ++i) tmp[i] = (i == c) ? 1.f : 0.f;
      lower_solve(tmp, tmp + D);
      upper_solve(tmp + D);
      for (int i = 0; i < D; ++i) Gi[i * D + c] = tmp[D + i];
    }
    for (int i = 0; i < D; ++i) dT[i] = 0.f;
    for (int n = 0; n < N; ++n) {
      const float* x = X + (size_t)n * D;
      float hq = 0.f;
      for (int i = 0; i < D; ++i) hq += x[i] * dotf(Gi + (size_t)i * D, x, D);
      const float u = dotf(x, v, D);
      const float sg = s[n], wp = sg * (1.f - sg) * (1.f - 2.f * sg);
      const float t = 0.5f * wp * (hq - u * u);
      for (int i = 0; i < D; ++i) dT[i] += t * x[i];
    }
  }
};

static void rmhmc_chain(const ocpu_problem& P, Key key, float* q0, float* l0, float* g0, Scratch& S, Out* o,
                        float* draw_out, float* prop_out) {
  const int D = P.D, N = P.N;
  size_t off = 0;
  LogReg tg{N, D, (float)P.prior_precision, P.X, P.y, S.take(N, off), S.take((size_t)D * D, off),
            S.take((size_t)D * D, off), S.take(N, off)};
  float* q = S.take(D, off); float* p = S.take(D, off); float* v = S.take(D, off); float* g = S.take(D, off);
  float* dT = S.take(D, off); float* qi = S.take(D, off); float* pi = S.take(D, off); float* qn = S.take(D, off);
  float* pn = S.take(D, off); float* z = S.take(D, off); float* tmp = S.take(2 * (size_t)D, off);
  const float he = 0.5f * (float)P.step_size;
  const Key k_m = split_index(key, 2, 0), k_a = split_index(key, 2, 1);
  normal(k_m, D, z);
  std::memcpy(q, q0, sizeof(float) * D);
  tg.sig_grad(q, g);
  float ld = tg.factor();
  for (int i = 0; i < D; ++i) {  // rmhmc/metrics.py:45-58: p = chol(G) z
    float t = 0.f;
    for (int j = 0; j <= i; ++j) t += tg.G[i * D + j] * z[j];
    p[i] = t;
    if (draw_out) draw_out[i] = t;
  }
  // kinetic energy rmhmc/metrics.py:60-74
  auto kinetic = [&](float logdiag) {
    tg.lower_solve(p, tmp);
    return 0.5f * dotf(tmp, tmp, D) + HALF_LOG_2PI * (float)D + logdiag;
  };
  const float H0 = -*l0 + kinetic(ld);
  // f(q, p) from (qi, pi): the map of rmhmc/integrators.py:119-142 at the current (q, p); s, g, factor are current
  auto map = [&](const float* qa, const float* pa) {
    tg.kinetic_grad(p, v, dT, tmp);
    for (int i = 0; i < D; ++i) {
      qn[i] = qa[i] + he * v[i];
      pn[i] = pa[i] - he * (dT[i] - g[i]);
    }
  };
  auto eval_at = [&](const float* qq) {
    tg.sig_grad(qq, g);
    return tg.factor();
  };
  int iters_total = 0;
  for (int step = 0; step < P.L; ++step) {
    std::memcpy(qi, q, sizeof(float) * D);
    std::memcpy(pi, p, sizeof(float) * D);
    if (step > 0) ld = eval_at(q);
    map(qi, pi);
    auto norm = [&]() {
      float m = 0.f;
      bool bad = false;
      for (int i = 0; i < D; ++i) {
        const float a = std::fabs(qn[i] - q[i]), b = std::fabs(pn[i] - p[i]);
        bad = bad || std::isnan(a) || std::isnan(b);
        m = std::fmax(m, std::fmax(a, b));
      }
      return bad ? (float)NAN : m;
    };
    float nrm = norm();
    std::memcpy(q, qn, sizeof(float) * D);
    std::memcpy(p, pn, sizeof(float) * D);
    int n = 0;
    while (n < P.fp_max_iters && std::isfinite(nrm) && nrm < (float)P.fp_div_tol && nrm > (float)P.fp_tol) {
      ld = eval_at(q);
      map(qi, pi);
      nrm = norm();
      std::memcpy(q, qn, sizeof(float) * D);
      std::memcpy(p, pn, sizeof(float) * D);
      ++n;
    }
    iters_total += n;
    // explicit update from the midpoint :147-148
    ld = eval_at(q);
    map(q, p);
    std::memcpy(q, qn, sizeof(float) * D);
    std::memcpy(p, pn, sizeof(float) * D);
  }
  ld = eval_at(q);  // end of trajectory :150-154
  const float l = tg.logp(q);
  for (int i = 0; i < D; ++i) p[i] = -p[i];  // flip_momentum rmhmc/rmhmc.py:443-462
  const float H1 = -l + kinetic(ld);
  metropolis(P, k_a, H0, H1, o);
  o->fp_iters = iters_total;
  if (prop_out) std::memcpy(prop_out, q, sizeof(float) * D);
  if (o->accepted) {
    std::memcpy(q0, q, sizeof(float) * D);
    std::memcpy(g0, g, sizeof(float) * D);
    *l0 = l;
  }
}

static size_t scratch_floats(const ocpu_problem& P) {
  const size_t D = P.D, N = P.N > 0 ? P.N : 0;
  return 16 * D + 2 * D * D + 2 * N + 64;
}

static int check(const ocpu_problem* p) {
  if (!p || p->D < 2 || p->L < 0) return -1;
  if (p->sampler == OCPU_RMHMC) return (p->target == OCPU_LOGREG && p->X && p->y && p->N > 0) ? 0 : -2;
  return p->target == OCPU_FUNNEL ? 0 : -2;  // lmc / lmcmonge: Neal's funnel
}

static void one_chain(const ocpu_problem& P, Key key, int64_t c, float* pos, float* logp, float* grad, float* vol,
                      Scratch& S, Out* o, const ocpu_info* info) {
  const int D = P.D;
  float* draw = (info && info->draw) ? info->draw + c * D : nullptr;
  float* prop = (info && info->proposal_position) ? info->proposal_position + c * D : nullptr;
  o->fp_iters = 0;
  if (P.sampler == OCPU_LMCMONGE) lmcmonge_chain(P, key, pos + c * D, logp + c, grad + c * D, vol + c, S, o, draw, prop);
  else if (P.sampler == OCPU_LMC) lmc_chain(P, key, pos + c * D, logp + c, grad + c * D, vol + c, S, o, draw, prop);
  else rmhmc_chain(P, key, pos + c * D, logp + c, grad + c * D, S, o, draw, prop);
  if (info) {
    if (info->acceptance_rate) info->acceptance_rate[c] = o->accept_rate;
    if (info->is_accepted) info->is_accepted[c] = o->accepted;
    if (info->energy) info->energy[c] = o->energy;
    if (info->initial_energy) info->initial_energy[c] = o->initial_energy;
    if (info->accept_uniform) info->accept_uniform[c] = o->u;
    if (info->fp_iters) info->fp_iters[c] = o->fp_iters;
  }
}

}  // namespace ocpu

using namespace ocpu;

extern "C" {

int ocpu_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ocpu_init(const ocpu_problem* p, int64_t C, const float* position, float* logdensity, float* logdensity_grad,
              float* volume_adjustment) {
  if (int rc = check(p)) return rc;
  const int D = p->D;
#pragma omp parallel num_threads(p->threads > 0 ? p->threads : ocpu_max_threads())
  {
    Scratch S;
    S.buf.resize(scratch_floats(*p));
#pragma omp for schedule(static)
    for (int64_t c = 0; c < C; ++c) {
      const float* q = position + c * D;
      if (p->target == OCPU_FUNNEL) {
        const Funnel tg{D, (float)p->sigma};
        logdensity[c] = tg.logp(q);
        tg.grad(q, logdensity_grad + c * D);
      } else {
        size_t off = 0;
        LogReg tg{p->N, D, (float)p->prior_precision, p->X, p->y, S.take(p->N, off), nullptr, nullptr, nullptr};
        logdensity[c] = tg.logp(q);
        tg.sig_grad(q, logdensity_grad + c * D);
      }
      if (volume_adjustment) volume_adjustment[c] = 0.f;
    }
  }
  return 0;
}

int ocpu_step(const ocpu_problem* p, int64_t C, const uint32_t* keys, float* position, float* logdensity,
              float* logdensity_grad, float* volume_adjustment, const ocpu_info* info) {
  if (int rc = check(p)) return rc;
  if (p->sampler != OCPU_RMHMC && !volume_adjustment) return -3;
#pragma omp parallel num_threads(p->threads > 0 ? p->threads : ocpu_max_threads())
  {
    Scratch S;
    S.buf.resize(scratch_floats(*p));
#pragma omp for schedule(dynamic, 1)
    for (int64_t c = 0; c < C; ++c) {
      Out o;
      one_chain(*p, Key{keys[2 * c], keys[2 * c + 1]}, c, position, logdensity, logdensity_grad, volume_adjustment, S, &o,
                info);
    }
  }
  return 0;
}

int ocpu_run(const ocpu_problem* p, int64_t C, const uint32_t* root_key, int64_t first_transition, int64_t T,
             int64_t total_transitions, int64_t chain_offset, int64_t total_chains, float* position, float* logdensity,
             float* logdensity_grad, float* volume_adjustment, double* mean_accept) {
  if (int rc = check(p)) return rc;
  if (p->sampler != OCPU_RMHMC && !volume_adjustment) return -3;
  double acc_sum = 0.0;
  std::vector<Key> tkeys((size_t)T);
  for (int64_t t = 0; t < T; ++t)
    tkeys[t] = split_index(Key{root_key[0], root_key[1]}, (uint32_t)total_transitions, (uint32_t)(first_transition + t));
#pragma omp parallel num_threads(p->threads > 0 ? p->threads : ocpu_max_threads()) reduction(+ : acc_sum)
  {
    Scratch S;
    S.buf.resize(scratch_floats(*p));
    // chains are independent: each thread carries its chains through all T transitions (state stays in cache)
#pragma omp for schedule(dynamic, 1)
    for (int64_t c = 0; c < C; ++c) {
      for (int64_t t = 0; t < T; ++t) {
        const Key k = split_index(tkeys[t], (uint32_t)total_chains, (uint32_t)(chain_offset + c));
        Out o;
        one_chain(*p, k, c, position, logdensity, logdensity_grad, volume_adjustment, S, &o, nullptr);
        acc_sum += o.accept_rate;
      }
    }
  }
  if (mean_accept) *mean_accept = acc_sum / (double)(C * T > 0 ? C * T : 1);
  return 0;
}

}  // extern "C"
