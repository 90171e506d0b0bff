// Cross-chain statistics of the ChEES adaptation (adaptation/chees_adaptation_riemanian.py:102-219,
// `compute_parameters`): everything that reduces over the chain axis, as plain SUMS so that shards on different GPUs
// combine with one all-reduce(sum) per pass.  The scalar recurrences (dual averaging on the harmonic-mean acceptance,
// the optimiser step on the log trajectory length, the moving averages) run on the host from these sums.
#include "launch.h"

namespace gb {

// pass 1: out[0, D) = sum_c proposal (NaN entries skipped, jnp.nanmean :158-163), [D, 2D) = their counts,
//         [2D, 3D) = sum_c initial, [3D, 4D) = counts, [4D] = sum over non-divergent chains of 1 / acceptance
//         (harmonic mean :145-147), [4D + 1] = number of non-divergent chains
__global__ void __launch_bounds__(256) k_chees_moments(const float* __restrict__ prop, const float* __restrict__ init,
                                                       const float* __restrict__ acc, const unsigned char* __restrict__ div,
                                                       long long C, int D, double* out) {
  const long long c0 = (long long)blockIdx.x * 64;
  const long long c1 = c0 + 64 < C ? c0 + 64 : C;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double sp = 0.0, si = 0.0;
    int np = 0, ni = 0;
    for (long long c = c0; c < c1; ++c) {
      const float p = prop[c * D + d], q = init[c * D + d];
      if (!isnan(p)) { sp += (double)p; ++np; }
      if (!isnan(q)) { si += (double)q; ++ni; }
    }
    atomicAdd(&out[d], sp);
    atomicAdd(&out[D + d], (double)np);
    atomicAdd(&out[2 * D + d], si);
    atomicAdd(&out[3 * D + d], (double)ni);
  }
  if (threadIdx.x >= 192) {  // the last two warps: acceptance statistics of the tile's 64 chains
    const long long c = c0 + (threadIdx.x - 192);
    double h = 0.0, n = 0.0;
    if (c < c1 && !div[c]) { h = 1.0 / (double)acc[c]; n = 1.0; }
    for (int o = 16; o > 0; o >>= 1) {
      h += __shfl_xor_sync(0xffffffffu, h, o);
      n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&out[4 * D], h);
      atomicAdd(&out[4 * D + 1], n);
    }
  }
}

// pass 2 (means = the all-reduced nanmeans, [0, D) proposals, [D, 2D) initials): per chain
//   g_c = (|p_c - pbar|^2 - |q_c - qbar|^2) ((p_c - pbar) . v_c)                        :176-184
// out[0] = sum over non-divergent chains of acceptance_c g_c, out[1] = sum of acceptance_c  :185-187
__global__ void __launch_bounds__(256) k_chees_gradient(const float* __restrict__ prop, const float* __restrict__ vel,
                                                        const float* __restrict__ init, const float* __restrict__ acc,
                                                        const unsigned char* __restrict__ div, const double* __restrict__ means,
                                                        long long C, int D, double* out) {
  __shared__ double red[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sg = 0.0, sa = 0.0;
  for (long long c = (long long)blockIdx.x * 8 + warp; c < C; c += (long long)gridDim.x * 8) {
    float pp = 0.f, qq = 0.f, pv = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float pc = prop[c * D + d] - (float)means[d];
      const float qc = init[c * D + d] - (float)means[D + d];
      pp = fmaf(pc, pc, pp);
      qq = fmaf(qc, qc, qq);
      pv = fmaf(pc, vel[c * D + d], pv);
    }
    for (int o = 16; o > 0; o >>= 1) {
      pp += __shfl_xor_sync(0xffffffffu, pp, o);
      qq += __shfl_xor_sync(0xffffffffu, qq, o);
      pv += __shfl_xor_sync(0xffffffffu, pv, o);
    }
    if (lane == 0 && !div[c]) {
      const float a = acc[c];
      sg += (double)(a * ((pp - qq) * pv));
      sa += (double)a;
    }
  }
  if (lane == 0) { red[0][warp] = sg; red[1][warp] = sa; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double g = 0.0, a = 0.0;
    for (int i = 0; i < 8; ++i) { g += red[0][i]; a += red[1][i]; }
    atomicAdd(&out[0], g);
    atomicAdd(&out[1], a);
  }
}

}  // namespace gb

using namespace gb;

extern "C" {

int gb200_chees_moments(const void* proposal_position, const void* initial_position, const void* acceptance_rate,
                        const uint8_t* is_divergent, int64_t C, int32_t D, double* out, void* stream) {
  if (!proposal_position || !initial_position || !acceptance_rate || !is_divergent || !out || C < 1 || D < 1) {
    set_error("chees_moments: bad argument");
    return GB200_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double) * (4 * (size_t)D + 2), s);
  k_chees_moments<<<(unsigned)((C + 63) / 64), 256, 0, s>>>((const float*)proposal_position, (const float*)initial_position,
                                                            (const float*)acceptance_rate, is_divergent, C, D, out);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

int gb200_chees_gradient(const void* proposal_position, const void* proposal_velocity, const void* initial_position,
                         const void* acceptance_rate, const uint8_t* is_divergent, const double* means, int64_t C, int32_t D,
                         double* out, void* stream) {
  if (!proposal_position || !proposal_velocity || !initial_position || !acceptance_rate || !is_divergent || !means || !out ||
      C < 1 || D < 1) {
    set_error("chees_gradient: bad argument");
    return GB200_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(double) * 2, s);
  long long blocks = (C + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_chees_gradient<<<(unsigned)blocks, 256, 0, s>>>((const float*)proposal_position, (const float*)proposal_velocity,
                                                    (const float*)initial_position, (const float*)acceptance_rate, is_divergent,
                                                    means, C, D, out);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

}  // extern "C"
