#include "lmcmonge.cuh"
#include "launch.h"

namespace gb {

template <typename R, class Target, bool ALLOW_EXACT>
static int launch_lmcmonge_t(const TransArgs& a, const Target& tg, LayoutChoice lay, cudaStream_t s) {
  int grid, block;
  launch_shape(a.C, lay.lpc, &grid, &block);
  const bool unit = a.inv_mass == nullptr;
  if constexpr (ALLOW_EXACT) {
    // lean instantiations (no Info, no overrides, no adaptation, legacy threefry, unit mass)
    const bool lean = unit && lean_launch(a);
#define GB_XL(E, L, HSV)                                                                    \
  if (lean && a.half_step == HSV && lay.epl == E && lay.lpc == L && a.D == E * L) {         \
    lmcmonge_kernel<R, Target, E, L, true, true, HSV, true><<<grid, block, (size_t)block * lay.epl * sizeof(R), s>>>(a, tg); \
    GB_CHECK_LAUNCH();                                                                      \
    return GB200_OK;                                                                        \
  }
#define GB_XL3(E, L) GB_XL(E, L, GB200_HALF_STEP_OMEGA) GB_XL(E, L, GB200_HALF_STEP_OMEGA_FIXED) GB_XL(E, L, GB200_HALF_STEP_OMEGATILDE)
    GB_MY_EXACT(GB_XL3)
#undef GB_XL3
#undef GB_XL
#define GB_XE(E, L)                                                                         \
  if (lay.epl == E && lay.lpc == L && a.D == E * L) {                                       \
    if (unit) lmcmonge_kernel<R, Target, E, L, true, true><<<grid, block, (size_t)block * lay.epl * sizeof(R), s>>>(a, tg);   \
    else lmcmonge_kernel<R, Target, E, L, true, false><<<grid, block, (size_t)block * lay.epl * sizeof(R), s>>>(a, tg);       \
    GB_CHECK_LAUNCH();                                                                      \
    return GB200_OK;                                                                        \
  }
  GB_MY_EXACT(GB_XE)
#undef GB_XE
  }
#define GB_X(E, L)                                                                    \
  if (lay.epl == E && lay.lpc == L) {                                                 \
    lmcmonge_kernel<R, Target, E, L, false, false><<<grid, block, (size_t)block * lay.epl * sizeof(R), s>>>(a, tg);     \
    GB_CHECK_LAUNCH();                                                                \
    return GB200_OK;                                                                  \
  }
  GB_MY_LAYOUTS(GB_X)
#undef GB_X
  set_error("lmcmonge: no kernel for layout (%d,%d)", lay.epl, lay.lpc);
  return GB200_ERR_UNSUPPORTED;
}

// float64 instantiations (jax_enable_x64 parity, north-star tolerance rel 1e-10): one lane per chain, three
// register-array sizes.  The normal / uniform draws of a float64 launch are the float32 streams widened
// (JAX's 64-bit draws are not restated); parity tests pass the oracle's draws as overrides.
#define GB_F64_LAYOUTS(X) X(2) X(8) X(20) X(32)
#if GB_LPC == 1
static int launch_lmcmonge_f64(const TransArgs& a, const Funnel<double>& tg, cudaStream_t s) {
  int grid, block;
  launch_shape(a.C, 1, &grid, &block);
#define GB_X(E)                                                                                                   \
  if (a.D <= E) {                                                                                                 \
    lmcmonge_kernel<double, Funnel<double>, E, 1, false, false><<<grid, block, (size_t)block * E * sizeof(double), s>>>(a, tg); \
    GB_CHECK_LAUNCH();                                                                                            \
    return GB200_OK;                                                                                              \
  }
  GB_F64_LAYOUTS(GB_X)
#undef GB_X
  set_error("lmcmonge: float64 is built for D <= 32");
  return GB200_ERR_UNSUPPORTED;
}
#endif

int GB_LPC_NAME(launch_lmcmonge)(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s) {
  if (dtype == GB200_F64) {
#if GB_LPC == 1
    if (t.kind == GB200_TARGET_FUNNEL) {
      Funnel<double> tg;
      tg.setup(t);
      return launch_lmcmonge_f64(a, tg, s);
    }
#endif
    set_error("lmcmonge: float64 is built for the funnel target with one lane per chain (D <= 32)");
    return GB200_ERR_UNSUPPORTED;
  }
  switch (t.kind) {
    case GB200_TARGET_FUNNEL: {
      Funnel<float> tg;
      tg.setup(t);
      return launch_lmcmonge_t<float, Funnel<float>, true>(a, tg, lay, s);
    }
    case GB200_TARGET_GAUSSIAN: {
      GaussianDiag<float> tg;
      tg.setup(t);
      return launch_lmcmonge_t<float, GaussianDiag<float>, false>(a, tg, lay, s);
    }
#if GB_LPC == 1
    case GB200_TARGET_BANANA: {
      Banana<float> tg;
      tg.setup(t);
      int grid, block;
      launch_shape(a.C, 1, &grid, &block);
      lmcmonge_kernel<float, Banana<float>, 2, 1, false, false><<<grid, block, (size_t)block * 2 * sizeof(float), s>>>(a, tg);
      GB_CHECK_LAUNCH();
      return GB200_OK;
    }
#endif
    default:
      set_error("lmcmonge: target kind %d has no in-kernel implementation", t.kind);
      return GB200_ERR_UNSUPPORTED;
  }
}

}  // namespace gb
