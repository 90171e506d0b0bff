#include "rmhmc.cuh"
#if GB_LPC == 1
#include "softabs.cuh"
#endif
#include "launch.h"

namespace gb {

template <typename R, class Target, class Metric, bool ALLOW_EXACT>
static int launch_rmhmc_t(const TransArgs& a, const Target& tg, LayoutChoice lay, cudaStream_t s) {
  int grid, block;
  launch_shape(a.C, lay.lpc, &grid, &block);
  if constexpr (ALLOW_EXACT) {
    // No lean instantiations for rmhmc: ptxas contracts a few FMAs differently once the Info stores are
    // gone (measured: fused != stepwise in the last bit), and bit-identity of fused and stepwise
    // launches is part of the contract.  lmc / lmcmonge lean kernels are bit-identical (tested).
#define GB_XE(E, L)                                                                   \
  if (lay.epl == E && lay.lpc == L && a.D == E * L) {                                 \
    rmhmc_kernel<R, Target, Metric, E, L, true><<<grid, block, (size_t)block * lay.epl * sizeof(R), s>>>(a, tg);          \
    GB_CHECK_LAUNCH();                                                                \
    return GB200_OK;                                                                  \
  }
    GB_MY_EXACT(GB_XE)
#undef GB_XE
  }
#define GB_X(E, L)                                                                    \
  if (lay.epl == E && lay.lpc == L) {                                                 \
    rmhmc_kernel<R, Target, Metric, E, L, false><<<grid, block, (size_t)block * lay.epl * sizeof(R), s>>>(a, tg);         \
    GB_CHECK_LAUNCH();                                                                \
    return GB200_OK;                                                                  \
  }
  GB_MY_LAYOUTS(GB_X)
#undef GB_X
  set_error("rmhmc: no kernel for layout (%d,%d)", lay.epl, lay.lpc);
  return GB200_ERR_UNSUPPORTED;
}

// float64 instantiations (jax_enable_x64 parity, north-star tolerance rel 1e-10): one lane per chain, three
// register-array sizes.  The normal / uniform draws of a float64 launch are the float32 streams widened
// (JAX's 64-bit draws are not restated); parity tests pass the oracle's draws as overrides.
#define GB_F64_LAYOUTS(X) X(2) X(8) X(20) X(32)
#if GB_LPC == 1
static int launch_rmhmc_f64(const TransArgs& a, const Funnel<double>& tg, cudaStream_t s) {
  int grid, block;
  launch_shape(a.C, 1, &grid, &block);
#define GB_X(E)                                                                                                   \
  if (a.D <= E) {                                                                                                 \
    rmhmc_kernel<double, Funnel<double>, FunnelArrowH<double>, E, 1, false><<<grid, block, (size_t)block * E * sizeof(double), s>>>(a, tg); \
    GB_CHECK_LAUNCH();                                                                                            \
    return GB200_OK;                                                                                              \
  }
  GB_F64_LAYOUTS(GB_X)
#undef GB_X
  set_error("rmhmc: float64 is built for D <= 32");
  return GB200_ERR_UNSUPPORTED;
}
#endif

int GB_LPC_NAME(launch_rmhmc)(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s) {
  if (dtype == GB200_F64) {
#if GB_LPC == 1
    if (t.kind == GB200_TARGET_FUNNEL && t.metric == GB200_METRIC_TARGET) {
      Funnel<double> tg;
      tg.setup(t);
      return launch_rmhmc_f64(a, tg, s);
    }
#endif
    set_error("rmhmc: float64 is built for the funnel target and metric with one lane per chain (D <= 32)");
    return GB200_ERR_UNSUPPORTED;
  }
  switch (t.kind) {
    case GB200_TARGET_FUNNEL: {
      if (t.metric == GB200_METRIC_SOFTABS) {
#if GB_LPC == 1
        if (t.D == 2 && lay.epl == 2) {
          FunnelSA<float> sa;
          sa.setup(t);
          int grid, block;
          launch_shape(a.C, 1, &grid, &block);
          rmhmc_kernel<float, FunnelSA<float>, SoftAbs2H<float>, 2, 1, true><<<grid, block, (size_t)block * 2 * sizeof(float), s>>>(a, sa);
          GB_CHECK_LAUNCH();
          return GB200_OK;
        }
#endif
        set_error("rmhmc: the SoftAbs metric is built for the D = 2 funnel with one lane per chain");
        return GB200_ERR_UNSUPPORTED;
      }
      Funnel<float> tg;
      tg.setup(t);
      if (t.metric == GB200_METRIC_IDENTITY)
        return launch_rmhmc_t<float, Funnel<float>, IdentityMetricH<float, Funnel<float>>, false>(a, tg, lay, s);
      return launch_rmhmc_t<float, Funnel<float>, FunnelArrowH<float>, true>(a, tg, lay, s);
    }
    case GB200_TARGET_GAUSSIAN: {
      GaussianDiag<float> tg;
      tg.setup(t);
      if (t.metric == GB200_METRIC_IDENTITY)
        return launch_rmhmc_t<float, GaussianDiag<float>, IdentityMetricH<float, GaussianDiag<float>>, false>(a, tg, lay, s);
      return launch_rmhmc_t<float, GaussianDiag<float>, TargetDiagMetricH<float, GaussianDiag<float>>, false>(a, tg, lay, s);
    }
#if GB_LPC == 1
    case GB200_TARGET_BANANA: {
      Banana<float> tg;
      tg.setup(t);
      int grid, block;
      launch_shape(a.C, 1, &grid, &block);
      rmhmc_kernel<float, Banana<float>, IdentityMetricH<float, Banana<float>>, 2, 1, false><<<grid, block, (size_t)block * 2 * sizeof(float), s>>>(a, tg);
      GB_CHECK_LAUNCH();
      return GB200_OK;
    }
#endif
    default:
      set_error("rmhmc: target kind %d has no in-kernel implementation", t.kind);
      return GB200_ERR_UNSUPPORTED;
  }
}

}  // namespace gb
