// CPU restatement (C++ / OpenMP) of geomjax's static Riemannian transition kernels -- ORACLE / CPU BASELINE ONLY.
// TEST INFRASTRUCTURE: used by tests/, bench.py's cpu_baseline / --impl reference legs; never by geomjax_b200/.
// The NumPy oracle (oracle/samplers.py) is the specification this is checked against (tests/test_oracle_cpp.py).
#pragma once
#include <stdint.h>

namespace ocpu {
struct Key {
  uint32_t a, b;
};
void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0, uint32_t* o1);
uint32_t bits_elem(Key k, uint32_t j, uint32_t n);
Key split_index(Key k, uint32_t num, uint32_t idx);  // jax.random.split(k, num)[idx]
float uniform01(Key k);
void normal(Key k, int n, float* out);  // jax.random.normal(k, (n,))
}  // namespace ocpu

extern "C" {

enum { OCPU_LMCMONGE = 0, OCPU_LMC = 1, OCPU_RMHMC = 2 };
enum { OCPU_FUNNEL = 0, OCPU_LOGREG = 1 };
enum { OCPU_OMEGA = 0, OCPU_OMEGA_FIXED = 1, OCPU_OMEGATILDE = 2 };

typedef struct ocpu_problem {
  int32_t sampler, target, D, L;       // L = num_integration_steps
  int32_t N;                           // logreg: data rows
  int32_t half_step;                   // lmcmonge variant
  int32_t fp_max_iters;                // rmhmc (100)
  int32_t threads;                     // OpenMP threads (0 = runtime default)
  double step_size, sigma, alpha2, prior_precision, divergence_threshold, fp_tol, fp_div_tol;
  const float* X;                      // logreg: [N, D] row-major
  const float* y;                      // logreg: [N]
  const float* inv_mass;               // lmcmonge: [D]
} ocpu_problem;

typedef struct ocpu_info {  // all optional (NULL = not wanted), one entry per chain
  float* draw;              // [C, D] momentum (rmhmc) / velocity (lmc, lmcmonge) drawn at the start
  float* acceptance_rate;   // [C]
  uint8_t* is_accepted;     // [C]
  float* energy;            // [C] proposal energy
  float* initial_energy;    // [C]
  float* proposal_position; // [C, D]
  float* accept_uniform;    // [C]
  int32_t* fp_iters;        // [C] rmhmc: fixed-point iterations summed over the L steps
} ocpu_info;

// State arrays are updated in place: position [C, D], logdensity [C], logdensity_grad [C, D], volume_adjustment [C]
// (NULL for rmhmc).  init: logdensity, gradient, volume = 0 (rmhmc/rmhmc.py:96-98, lmcmc/lmc.py:98-101).
int ocpu_init(const ocpu_problem* p, int64_t C, const float* position, float* logdensity, float* logdensity_grad,
              float* volume_adjustment);
// One transition per chain with explicit per-chain keys [C, 2] == jax.vmap(kernel)(keys, states).
int ocpu_step(const ocpu_problem* p, int64_t C, const uint32_t* keys, float* position, float* logdensity,
              float* logdensity_grad, float* volume_adjustment, const ocpu_info* info);
// T transitions with the example's key tree split(split(root, total_transitions)[t], total_chains)[chain_offset + c]
// (examples/funnel/main.py:18,22); returns the mean acceptance rate through *mean_accept (may be NULL).
int ocpu_run(const ocpu_problem* p, int64_t C, const uint32_t* root_key, int64_t first_transition, int64_t T,
             int64_t total_transitions, int64_t chain_offset, int64_t total_chains, float* position, float* logdensity,
             float* logdensity_grad, float* volume_adjustment, double* mean_accept);
int ocpu_max_threads(void);
}
