/* geomb200.h -- C ABI of the B200-native batched-chain engine for geomjax's static
 * Riemannian transition kernels (rmhmc, lmc, lmcmonge).
 *
 * The reference (williwilliams3/geomjax) has no FFI layer: its operator API is Python
 *   geomjax.rmhmc / lmc / lmcmonge (logdensity_fn, step_size, metric_fn | inverse_mass_matrix,
 *                                   num_integration_steps, ...) -> SamplingAlgorithm(init, step)
 *   (geomjax/rmhmc/rmhmc.py:286-311, geomjax/lmcmc/lmc.py:321-346, geomjax/lmcmonge/lmc.py:378-405)
 * and users batch it with jax.vmap(kernel)(keys, states) (examples/funnel/main.py:13,19).
 * Each entry point below is what a typed-FFI custom call for that path would bind; the
 * reference interface it replaces is cited on every declaration.  INTEGRATION.md shows the
 * jax.ffi and ctypes stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; every buffer is caller-owned DEVICE memory,
 *     row-major (C, D) like the vmapped JAX arrays, float32 (dtype GB200_F32) unless stated;
 *   - every call is asynchronous on the caller's cudaStream_t (passed as void*), never
 *     synchronises and never allocates;
 *   - returns 0 on success or a negative gb200_status; the message is available through
 *     gb200_last_error() (thread-local).  Numerical failures (NaN energy, divergence,
 *     fixed-point non-convergence) are DATA reported in gb200_info, never error codes
 *     (mcmc/proposal.py:107-108, rmhmc/rmhmc.py:424);
 *   - re-entrant, no global mutable state besides the thread-local error string.
 */
#ifndef GEOMB200_H
#define GEOMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB200_VERSION 100

typedef enum gb200_status {
  GB200_OK = 0,
  GB200_ERR_INVALID_ARGUMENT = -1,
  GB200_ERR_UNSUPPORTED = -2,
  GB200_ERR_CUDA = -3
} gb200_status;

typedef enum gb200_dtype { GB200_F32 = 0, GB200_F64 = 1 } gb200_dtype;

/* jax_threefry_partitionable=False ("legacy", the reference's effective JAX range) or True. */
typedef enum gb200_threefry_mode { GB200_THREEFRY_LEGACY = 0, GB200_THREEFRY_PARTITIONABLE = 1 } gb200_threefry_mode;

typedef enum gb200_sampler { GB200_RMHMC = 0, GB200_LMC = 1, GB200_LMCMONGE = 2 } gb200_sampler;

/* lmcmonge/integrators.py:158-194 as written ("omega", default), the same with the
 * missing alpha2 restored ("omega_fixed"), and :197-230 ("omegatilde"). */
typedef enum gb200_half_step { GB200_HALF_STEP_OMEGA = 0, GB200_HALF_STEP_OMEGA_FIXED = 1, GB200_HALF_STEP_OMEGATILDE = 2 } gb200_half_step;

/* Built-in targets: replace the reference's logdensity_fn / metric_fn callables
 * (examples/funnel/main.py:28-54 is the only instance in the reference tree). */
typedef enum gb200_target_kind {
  GB200_TARGET_FUNNEL = 0,   /* Neal's funnel, pull-back ("fisher_metric_fn") metric; params[0]=sigma */
  GB200_TARGET_GAUSSIAN = 1, /* diagonal Gaussian; vec0=mean[D], vec1=precision[D]; metric=diag(precision) */
  GB200_TARGET_BANANA = 2,   /* D=2; params[0]=sigma1^2, params[1]=b; identity metric */
  GB200_TARGET_LOGREG = 3    /* Bayesian logistic regression; X[N,D] row-major, y[N]; params[0]=prior precision; Fisher+prior metric */
} gb200_target_kind;

typedef enum gb200_metric_kind {
  GB200_METRIC_TARGET = 0,   /* the target's own Riemannian metric */
  GB200_METRIC_IDENTITY = 1, /* metric_fn = lambda x: eye(D) (tests/test_samplers.py:25,37) */
  GB200_METRIC_SOFTABS = 2   /* SoftAbs of the Hessian (Betancourt 2013), params[7] = softabs alpha; rmhmc, funnel D = 2
                              * (the metric BASELINE.json configs[0] names; NEW, not in the reference) */
} gb200_metric_kind;

typedef struct gb200_target_desc {
  int32_t kind;          /* gb200_target_kind */
  int32_t metric;        /* gb200_metric_kind */
  int32_t D;
  int32_t reserved;
  int64_t N;             /* data rows (logreg) */
  double params[8];
  const void* X;         /* borrowed device pointers; must outlive every call using the descriptor */
  const void* y;
  const void* vec0;
  const void* vec1;
} gb200_target_desc;

/* kwargs of the reference kernels: rmhmc/rmhmc.py:286-311 (+ solver kwargs of
 * rmhmc/integrators.py:53-61), lmcmc/lmc.py:321-346, lmcmonge/lmc.py:378-405. */
typedef struct gb200_kernel_params {
  double step_size;                  /* scalar step size, used when step_size_per_chain == NULL */
  const void* step_size_per_chain;   /* optional [C] (vmapped adaptation); dtype of the call */
  int32_t num_integration_steps;
  int32_t threefry_mode;             /* gb200_threefry_mode */
  double divergence_threshold;       /* default 1000 */
  double fp_convergence_tol;         /* rmhmc: default 1e-6 */
  double fp_divergence_tol;          /* rmhmc: default 1e10 */
  int32_t fp_max_iters;              /* rmhmc: default 100 */
  int32_t half_step;                 /* lmcmonge: gb200_half_step */
  double alpha2;                     /* lmcmonge: default 1e-3 */
  const void* inverse_mass_matrix;   /* lmcmonge: [D] diagonal, NULL = ones */
  int32_t dtype;                     /* gb200_dtype */
  int32_t lanes_per_chain;           /* 0 = auto; else 1,2,4,8,16,32 (tuning knob) */
  int32_t inverse_mass_per_chain;    /* lmcmonge: 0 = inverse_mass_matrix is [D]; 1 = [C, D] (vmapped window adaptation) */
  int32_t reserved;
  const int32_t* num_integration_steps_per_chain; /* optional [C] (device): the dynamic kernels' per-chain draw of
                                        integration_steps_fn (rmhmc/rmhmc.py:179-244, lmcmc/lmc.py:185-252);
                                        num_integration_steps must then be an upper bound of its entries */
} gb200_kernel_params;

/* RMHMCState / LMCState (rmhmc/rmhmc.py:30-41, lmcmc/lmc.py:30-42, lmcmonge/lmc.py:32-44). */
typedef struct gb200_state {
  void* position;          /* [C, D] */
  void* logdensity;        /* [C] */
  void* logdensity_grad;   /* [C, D] */
  void* volume_adjustment; /* [C]; LMC kernels only, NULL for rmhmc */
} gb200_state;

/* RMHMCInfo / LMCInfo (rmhmc/rmhmc.py:58-93, lmcmc/lmc.py:60-95, lmcmonge/lmc.py:63-98).
 * Every field may be NULL (not written). */
typedef struct gb200_info {
  void* momentum;            /* [C, D] initial draw: momentum (rmhmc) or velocity (lmc, lmcmonge) */
  void* acceptance_rate;     /* [C] */
  uint8_t* is_accepted;      /* [C] */
  uint8_t* is_divergent;     /* [C] */
  void* energy;              /* [C] energy of the proposed state */
  void* proposal_position;   /* [C, D] info.proposal.state.position */
  void* proposal_momentum;   /* [C, D] (flipped) */
  void* proposal_velocity;   /* [C, D] (flipped) */
  void* proposal_logdensity; /* [C] */
  void* proposal_logdensity_grad; /* [C, D] */
  void* proposal_volume_adjustment; /* [C] */
  void* proposal_weight;     /* [C] H0 - H1 (NaN -> -inf) */
  void* initial_energy;      /* [C] H0 (extra; not in the reference Info) */
  int32_t* fp_iters;         /* [C] rmhmc: total fixed-point iterations over the trajectory (extra) */
  void* accept_uniform;      /* [C] the uniform the accept test used (extra, test surface) */
  void* noise;               /* [C, D] the standard normal draw z (extra, test surface) */
} gb200_info;

/* Where the per-chain keys of a transition come from.
 *   keys != NULL : explicit per-chain keys [C, 2] uint32 == vmap(kernel)(keys, states);
 *                  num_transitions must be 1.
 *   keys == NULL : derived in-kernel exactly like the driver loop of
 *                  examples/funnel/main.py:18,22: k = split(split(root, total_transitions)[t], total_chains)[chain_offset + c]
 *                  for t = first_transition .. first_transition + num_transitions - 1 (fused launch). */
typedef struct gb200_key_source {
  const uint32_t* keys;
  uint32_t root_key[2];
  int64_t first_transition;
  int64_t num_transitions;
  int64_t total_transitions;
  int64_t chain_offset;
  int64_t total_chains;
} gb200_key_source;

/* Optional per-launch extras (all may be NULL / zero). */
typedef struct gb200_run_opts {
  void* samples;               /* [num_transitions, C, D] positions after each transition (examples/funnel/main.py:20) */
  void* sample_accept;         /* [num_transitions, C] acceptance_rate per transition */
  const void* noise_override;  /* [C, D] use this z instead of drawing (test surface, num_transitions==1) */
  const void* uniform_override;/* [C] use this accept uniform (test surface, num_transitions==1) */
  void* dual_averaging;        /* [C, 5] (log_x, log_x_avg, step, avg_error, mu) IN THE STATE DTYPE: fused per-chain dual
                                  averaging (optimizers/dual_averaging.py:101-123); step size = exp(log_x).  samples,
                                  sample_accept and accept_sum are in the state dtype as well. */
  double da_target;            /* target acceptance rate (adaptation/step_size_adaptation.py:103) */
  double da_t0, da_gamma, da_kappa; /* (10, 0.05, 0.75) */
  void* workspace;             /* optional device scratch (>= 16 bytes): dynamic chain hand-out of the CTA-per-chain
                                  logistic-regression kernels (used when plan == NULL) */
  int64_t workspace_bytes;
  void* plan;                  /* gb200_plan* (below): rmhmc on the logistic-regression target runs the lock-step
                                  tcgen05 sampler */
  void* accept_sum;            /* [C] running sum: += acceptance_rate of every transition of the launch (caller zeroes);
                                  the per-chain mean acceptance of a fused run without a [T, C] buffer */
} gb200_run_opts;

int gb200_version(void);
/* Number of CUDA kernels this library has enqueued in this process (monotonic; bench.py's gpu_launches). */
long long gb200_kernel_launches(void);
const char* gb200_last_error(void);

/* ---- PRNG test surface: jax.random.split / bits / uniform / normal --------------------- */
/* split(keys[n], num) -> out[n, num, 2]            (rmhmc/rmhmc.py:158 etc.) */
int gb200_threefry_split(const uint32_t* keys, uint32_t* out, int64_t n_keys, int32_t num, int32_t mode, void* stream);
/* random_bits(keys[n], (count,)) -> out[n, count] */
int gb200_random_bits(const uint32_t* keys, uint32_t* out, int64_t n_keys, int32_t count, int32_t mode, void* stream);
/* uniform(keys[n], (count,), float32) -> out[n, count]   (mcmc/proposal.py:178 via bernoulli) */
int gb200_uniform_f32(const uint32_t* keys, float* out, int64_t n_keys, int32_t count, int32_t mode, void* stream);
/* normal(keys[n], (count,), float32) -> out[n, count]    (util.py:81-82) */
int gb200_normal_f32(const uint32_t* keys, float* out, int64_t n_keys, int32_t count, int32_t mode, void* stream);
/* split(split(root, total_transitions)[t], total_chains)[chain_offset + c], c < C -> out[C, 2] */
int gb200_chain_keys(const uint32_t root_key[2], int64_t t, int64_t total_transitions, int64_t chain_offset,
                     int64_t total_chains, uint32_t* out, int64_t C, int32_t mode, void* stream);

/* ---- init: X.init(position, logdensity_fn) ----------------------------------------------
 * rmhmc/rmhmc.py:96-98, lmcmc/lmc.py:98-101, lmcmonge/lmc.py:101-109.
 * Fills logdensity, logdensity_grad (and volume_adjustment = 0 when non-NULL) from position. */
int gb200_init(const gb200_target_desc* target, gb200_state state, int64_t C, int32_t dtype, void* stream);

/* ---- step: vmap(kernel)(keys, states) ---------------------------------------------------
 * rmhmc/rmhmc.py:131-174 (+416-462), lmcmc/lmc.py:135-180 (+451-499), lmcmonge/lmc.py:151-235 (+512-565).
 * state_in and state_out may alias field by field (in-place) or be disjoint. */
int gb200_step(int32_t sampler, const gb200_kernel_params* params, const gb200_target_desc* target,
               const gb200_key_source* keys, gb200_state state_in, gb200_state state_out,
               const gb200_info* info, const gb200_run_opts* opts, int64_t C, void* stream);

int gb200_rmhmc_step(const gb200_kernel_params* params, const gb200_target_desc* target, const gb200_key_source* keys,
                     gb200_state state_in, gb200_state state_out, const gb200_info* info, const gb200_run_opts* opts,
                     int64_t C, void* stream);
int gb200_lmc_step(const gb200_kernel_params* params, const gb200_target_desc* target, const gb200_key_source* keys,
                   gb200_state state_in, gb200_state state_out, const gb200_info* info, const gb200_run_opts* opts,
                   int64_t C, void* stream);
int gb200_lmcmonge_step(const gb200_kernel_params* params, const gb200_target_desc* target, const gb200_key_source* keys,
                        gb200_state state_in, gb200_state state_out, const gb200_info* info, const gb200_run_opts* opts,
                        int64_t C, void* stream);

/* ---- dual averaging: optimizers/dual_averaging.py:87-127 -------------------------------- */
/* da[C,5] <- init(step_size0[C]) */
int gb200_dual_averaging_init(void* da, const void* step_size0, int64_t C, int32_t dtype, void* stream);
/* da <- update(da, target - acceptance_rate[C]) */
int gb200_dual_averaging_update(void* da, const void* acceptance_rate, double target, double t0, double gamma,
                                double kappa, int64_t C, int32_t dtype, void* stream);

/* ---- diagnostics: geomjax/diagnostics.py:25-75 (rhat), :78-209 (ess) ---------------------
 * samples[T, C, D] (sample_axis=0, chain_axis=1 as in examples/funnel/main.py:77-78).
 * *_partial writes per-shard sufficient statistics (float64) that are SUM-all-reducible across
 * ranks; *_finalize turns the (reduced) statistics into rhat[D] / ess[D] on the host. */
/* stats[3*D+1]: sum_c mean_c, sum_c mean_c^2, sum_c var_c (ddof=1), then C */
int gb200_rhat_partial(const void* samples, int64_t T, int64_t C, int32_t D, double* stats, int32_t dtype, void* stream);
int gb200_rhat_finalize(const double* stats_host, int64_t T, int32_t D, double* rhat_host);
/* acov[num_lags*D]: sum_c autocovariance_c(lag) (biased, /T) for lag < num_lags */
int gb200_ess_partial(const void* samples, int64_t T, int64_t C, int32_t D, int32_t num_lags, double* acov,
                      int32_t dtype, void* stream);
/* returns per dim ess; truncated[d]=1 if Geyer's sequence was still positive at num_lags (ask for more lags) */
int gb200_ess_finalize(const double* acov_host, const double* rhat_stats_host, int64_t T, int64_t C_total, int32_t D,
                       int32_t num_lags, double* ess_host, uint8_t* truncated_host);

/* ---- streaming R-hat / ESS (geomjax/diagnostics.py:25-209 without the [T, C, D] sample tensor) ----------------
 * Samples are fed in blocks [Tb, C, D] (e.g. the sample buffer of one fused launch, reused); the workspace keeps per
 * series the shift, sum, sum of squares and the first / last `max_lags` samples, and per (lag, dim) the chain-summed
 * lagged products.  gb200_stream_diag_partial then yields exactly the statistics of gb200_rhat_partial (stats) and
 * gb200_ess_partial (acov, num_lags <= max_lags rounded up to 8): all-reduce them over ranks and finalise as usual.
 * T_prev = samples already fed (0 on the first block: resets the accumulators); T = total fed. */
int64_t gb200_stream_diag_workspace(int64_t C, int32_t D, int32_t max_lags);
int gb200_stream_diag_update(void* workspace, const void* samples, int64_t Tb, int64_t C, int32_t D, int32_t max_lags,
                             int64_t T_prev, int32_t dtype, void* stream);
int gb200_stream_diag_partial(void* workspace, int64_t T, int64_t C, int32_t D, int32_t max_lags, int32_t num_lags,
                              double* stats, double* acov, void* stream);

/* ---- batched metric evaluation: vmap(metric_fn)(position) -------------------------------------
 * metric_fn of the logistic-regression target (the reference calls metric_fn at every kinetic-energy /
 * velocity evaluation: rmhmc/metrics.py:46,62,121): G_c = X^T diag(s(1-s)) X + alpha I for every chain,
 * computed as ONE tcgen05 (3xTF32) GEMM over the chain dimension.  position [C, D] -> metric [C, D, D].
 * workspace: gb200_logreg_fisher_metric_workspace(target, C) bytes of device memory. */
int gb200_logreg_fisher_metric(const gb200_target_desc* target, const void* position, void* metric, void* workspace,
                               int64_t workspace_bytes, int64_t C, int32_t dtype, void* stream);
int64_t gb200_logreg_fisher_metric_workspace(const gb200_target_desc* target, int64_t C);
/* h[c, n] = x_n^T A_c x_n for per-chain symmetric matrices A[C, D, D] (e.g. A = G^-1: the h_n of rmhmc's
 * dT/dq_i = 1/2 sum_n w'_n x_ni (h_n - u_n^2), SURVEY Appendix B.1 GEMM 5) as ONE tcgen05 (3xTF32) GEMM
 * h[N, C] = Z[N, P] . vecsym(A)[P, C] over the chain dimension.  h is [C, ldh] float32, ldh >= N. */
int gb200_logreg_quadform(const gb200_target_desc* target, const void* matrices, void* h, int64_t ldh, void* workspace,
                          int64_t workspace_bytes, int64_t C, int32_t dtype, void* stream);
int64_t gb200_logreg_quadform_workspace(const gb200_target_desc* target, int64_t C);
/* ---- rmhmc on the logistic-regression target: the lock-step "rolling batch" sampler ---------------------------
 * rmhmc/rmhmc.py:131-174 (+416-462), rmhmc/integrators.py:53-156 under vmap.  Every evaluation of the
 * implicit-midpoint map runs for ALL unfinished chains at once with both D^2 N products on the tcgen05 GEMMs
 * above; each chain carries its own (transition, step, fixed-point iteration) state, so chains that converge early
 * move on (to their next step, or their next transition of a fused launch) instead of idling behind the slowest
 * chain as a vmapped while_loop makes them.  The round is the body of a CUDA-graph WHILE node: gb200_step stays
 * asynchronous.  A plan owns the graph and borrows caller memory:
 *   workspace: >= gb200_rmhmc_logreg_plan_workspace(target, C) bytes, 256-byte aligned, alive until plan_destroy;
 *   loop_mode: 0 = device-side WHILE node (falls back to 1 if the driver refuses it; see gb200_plan_loop_mode),
 *              1 = host-sequenced rounds (gb200_step then blocks the calling thread on its own stream).
 * Pass the plan in gb200_run_opts.plan to gb200_step / gb200_rmhmc_step (C <= the plan's C, same target). */
typedef struct gb200_plan gb200_plan;
int64_t gb200_rmhmc_logreg_plan_workspace(const gb200_target_desc* target, int64_t C);
int gb200_rmhmc_logreg_plan_create(const gb200_target_desc* target, int64_t C, void* workspace, int64_t workspace_bytes,
                                   int32_t loop_mode, void* stream, gb200_plan** out);
int gb200_plan_destroy(gb200_plan* plan);
const char* gb200_plan_loop_mode(const gb200_plan* plan);
/* rounds and chain-evaluations of the plan's last launch (synchronises the stream; measurement only) */
int gb200_plan_stats(const gb200_plan* plan, int64_t* rounds, int64_t* chain_evals, void* stream);
/* ONE round for explicit inputs (test surface; the unit the sampler's loop repeats), all arrays [C, D] / [C] float32:
 *   mode 0: the implicit-midpoint map (rmhmc/integrators.py:119-142) at (q, p) from (qi, pi):
 *           qn = qi + h dH/dp, pn = pi - h dH/dq, h = half_step; velocity = dH/dp = G(q)^-1 p, logdet = log det G(q),
 *           dHdq = dT/dq - grad logp;
 *   mode 1: end-of-trajectory state (rmhmc/integrators.py:150-154): logdensity, logdensity_grad, velocity, logdet;
 *   mode 2: mode 0 with the momentum drawn first: p holds z, p_out = chol(G(q)) z (rmhmc/metrics.py:45-58), pi = p_out.
 * Outputs may be NULL. */
int gb200_logreg_lockstep_eval(gb200_plan* plan, int32_t mode, const void* q, const void* p, const void* qi, const void* pi,
                               double half_step, void* qn, void* pn, void* p_out, void* logdensity, void* logdensity_grad,
                               void* velocity, void* logdet, void* dHdq, int64_t C, void* stream);

/* ---- ChEES adaptation: the cross-chain sums of compute_parameters (adaptation/chees_adaptation_riemanian.py:102-219),
 * float32 chains.  moments -> out[4 D + 2] = sum / count of the non-NaN proposal and initial positions per dimension,
 * sum of 1 / acceptance and count over the non-divergent chains; gradient (means[2 D] = the global nanmeans, device)
 * -> out[2] = sum of acceptance_c g_c and of acceptance_c over the non-divergent chains.  All plain sums: one
 * all-reduce(sum) per pass combines shards. */
int gb200_chees_moments(const void* proposal_position, const void* initial_position, const void* acceptance_rate,
                        const uint8_t* is_divergent, int64_t C, int32_t D, double* out, void* stream);
int gb200_chees_gradient(const void* proposal_position, const void* proposal_velocity, const void* initial_position,
                         const void* acceptance_rate, const uint8_t* is_divergent, const double* means, int64_t C, int32_t D,
                         double* out, void* stream);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* Runs a dependent-FMA microbenchmark (iters FFMA per thread on grid x block threads) for the FP32
 * roofline denominator; out[0] receives a checksum so the work cannot be elided. */
int gb200_fp32_peak_kernel(float* out, int32_t grid, int32_t block, int64_t iters, void* stream);
/* Algorithmic FP32 flop count per chain per integrator step used for roofline.achieved. */
double gb200_flops_per_chain_step(int32_t sampler, const gb200_target_desc* target);
/* Algorithmic FP32 flop count per chain per transition outside the integrator steps (normal transform,
 * draw, energies, prologue); bench.py amortises it over num_integration_steps. */
double gb200_flops_per_transition(int32_t sampler, const gb200_target_desc* target);

#ifdef __cplusplus
}
#endif
#endif /* GEOMB200_H */
