// X.init(position, logdensity_fn): logdensity and gradient from positions
// (rmhmc/rmhmc.py:96-98, lmcmc/lmc.py:98-101, lmcmonge/lmc.py:101-109).
#include "launch.h"

namespace gb {

template <typename R, class Target, int EPL, int LPC>
__global__ void __launch_bounds__(128) init_kernel(const Target tg, gb200_state st, long long C, int D) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long chain = tid / LPC;
  const bool active = chain < C;
  if (!active) chain = C - 1;
  using LAY = Lay<EPL, LPC, false>;
  LAY lay{D, (int)(tid % LPC)};
  R q[EPL], g[EPL];
  load_vec(lay, st.position, chain, q);
  typename Target::Ctx ctx = tg.prepare(lay, q);
  tg.grad(lay, ctx, q, g);
  if (active) {
    store_vec(lay, st.logdensity_grad, chain, g);
    if (lay.g == 0) {
      store_scalar<R>(st.logdensity, chain, tg.logp(ctx));
      store_scalar<R>(st.volume_adjustment, chain, R(0));
    }
  }
}

template <typename R, class Target>
static int launch_init_t(const Target& tg, gb200_state st, long long C, int D, LayoutChoice lay, cudaStream_t s) {
  int grid, block;
  launch_shape(C, lay.lpc, &grid, &block);
#define GB_X(E, L)                                                          \
  if (lay.epl == E && lay.lpc == L) {                                       \
    init_kernel<R, Target, E, L><<<grid, block, 0, s>>>(tg, st, C, D);      \
    GB_CHECK_LAUNCH();                                                      \
    return GB200_OK;                                                        \
  }
  GB_MY_LAYOUTS(GB_X)
#undef GB_X
  set_error("init: no kernel for layout (%d,%d)", lay.epl, lay.lpc);
  return GB200_ERR_UNSUPPORTED;
}

int GB_LPC_NAME(launch_init)(const gb200_target_desc& t, gb200_state st, long long C, LayoutChoice lay, int dtype, cudaStream_t s) {
  if (dtype == GB200_F64) {  // float64 (jax_enable_x64): funnel target, one lane per chain
#if GB_LPC == 1
    if (t.kind == GB200_TARGET_FUNNEL) {
      Funnel<double> tg;
      tg.setup(t);
      return launch_init_t<double>(tg, st, C, t.D, lay, s);
    }
#endif
    set_error("init: float64 is built for the funnel target with one lane per chain (D <= 32)");
    return GB200_ERR_UNSUPPORTED;
  }
  if (dtype != GB200_F32) {
    set_error("init: unknown dtype %d", dtype);
    return GB200_ERR_UNSUPPORTED;
  }
  switch (t.kind) {
    case GB200_TARGET_FUNNEL: {
      Funnel<float> tg;
      tg.setup(t);
      return launch_init_t<float>(tg, st, C, t.D, lay, s);
    }
    case GB200_TARGET_GAUSSIAN: {
      GaussianDiag<float> tg;
      tg.setup(t);
      return launch_init_t<float>(tg, st, C, t.D, lay, s);
    }
    case GB200_TARGET_BANANA: {
      Banana<float> tg;
      tg.setup(t);
      return launch_init_t<float>(tg, st, C, t.D, lay, s);
    }
    default:
      set_error("init: target kind %d has no in-kernel implementation", t.kind);
      return GB200_ERR_UNSUPPORTED;
  }
}

}  // namespace gb
