// Host-side launcher declarations (one translation unit per sampler so nvcc runs in parallel).
#pragma once
#include "dispatch.cuh"

namespace gb {
void set_error(const char* fmt, ...);
void count_launch();  // one kernel of this library enqueued (gb200_kernel_launches)
#define GB_DECL_LPC(n)                                                                                              \
  int launch_lmcmonge_lpc##n(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s); \
  int launch_lmc_lpc##n(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s);      \
  int launch_rmhmc_lpc##n(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s);    \
  int launch_init_lpc##n(const gb200_target_desc& t, gb200_state st, long long C, LayoutChoice lay, int dtype, cudaStream_t s);
GB_DECL_LPC(1) GB_DECL_LPC(2) GB_DECL_LPC(4) GB_DECL_LPC(8) GB_DECL_LPC(32)
#undef GB_DECL_LPC

// A launch is "lean" when it carries no Info outputs, no overrides, no adaptation and runs legacy
// threefry: the fused multi-transition launches of a sampling run.  Lean kernel instantiations
// compile all of that out (the fused kernels are instruction-cache bound).
inline bool lean_launch(const TransArgs& a) {
  const gb200_info& f = a.info;
  return a.mode == GB200_THREEFRY_LEGACY && a.opts.noise_override == nullptr && a.opts.uniform_override == nullptr &&
         a.opts.dual_averaging == nullptr && a.step_size_per_chain == nullptr && a.steps_per_chain == nullptr &&
         !f.noise && !f.momentum &&
         !f.acceptance_rate && !f.is_accepted && !f.is_divergent && !f.energy && !f.proposal_position &&
         !f.proposal_velocity && !f.proposal_momentum && !f.proposal_logdensity && !f.proposal_logdensity_grad &&
         !f.proposal_volume_adjustment && !f.proposal_weight && !f.initial_energy && !f.accept_uniform && !f.fp_iters;
}

int launch_rmhmc_logreg(const TransArgs& a, const gb200_target_desc& t, int dtype, cudaStream_t s);
int launch_rmhmc_lockstep(const TransArgs& a, const gb200_target_desc& t, gb200_plan* plan, int dtype, cudaStream_t s);
int launch_init_logreg(const gb200_target_desc& t, gb200_state st, long long C, int dtype, cudaStream_t s);

#define GB_BY_LPC(base, lay, ...)                          \
  switch ((lay).lpc) {                                     \
    case 1: return base##_lpc1(__VA_ARGS__);               \
    case 2: return base##_lpc2(__VA_ARGS__);               \
    case 4: return base##_lpc4(__VA_ARGS__);               \
    case 8: return base##_lpc8(__VA_ARGS__);               \
    case 32: return base##_lpc32(__VA_ARGS__);             \
    default: set_error("no kernels for lanes_per_chain=%d", (lay).lpc); return GB200_ERR_UNSUPPORTED; \
  }
inline int launch_lmcmonge(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s) { GB_BY_LPC(launch_lmcmonge, lay, a, t, lay, dtype, s) }
inline int launch_lmc(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s) { GB_BY_LPC(launch_lmc, lay, a, t, lay, dtype, s) }
inline int launch_rmhmc(const TransArgs& a, const gb200_target_desc& t, LayoutChoice lay, int dtype, cudaStream_t s) { GB_BY_LPC(launch_rmhmc, lay, a, t, lay, dtype, s) }
inline int launch_init(const gb200_target_desc& t, gb200_state st, long long C, LayoutChoice lay, int dtype, cudaStream_t s) { GB_BY_LPC(launch_init, lay, t, st, C, lay, dtype, s) }
}  // namespace gb

#define GB_CHECK_LAUNCH()                                                        \
  do {                                                                           \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) {                                                     \
      gb::set_error("CUDA launch failed: %s", cudaGetErrorString(e_));           \
      return GB200_ERR_CUDA;                                                     \
    }                                                                            \
    gb::count_launch();                                                          \
  } while (0)
