// R-hat and ESS as shard-local sufficient statistics + host finalisation.
// Reference: geomjax/diagnostics.py:25-75 (potential_scale_reduction), :78-209
// (effective_sample_size).  samples[T, C, D] with sample_axis=0, chain_axis=1
// (examples/funnel/main.py:77-78).  Every partial is a plain SUM over chains so that shards on
// different GPUs combine with one all-reduce(sum).
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "launch.h"

namespace gb {

// ---- R-hat partial: per (chain, dim) mean and ddof=1 variance over samples, summed over chains.
// stats layout: [0,D) sum_c mean; [D,2D) sum_c mean^2; [2D,3D) sum_c var; [3D] chains counted.
template <typename R>
__global__ void k_rhat_partial(const R* __restrict__ x, long long T, long long C, int D, double* stats) {
  extern __shared__ double sh[];  // 3*D
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const long long CD = C * D;
  // stride is a multiple of D when gridDim*blockDim is; enforce by rounding the stride
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long stride = (nthreads / D) * D;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (stride > 0 && tid < stride) {
    const int d = (int)(tid % D);
    double sm = 0.0, sm2 = 0.0, sv = 0.0;
    for (long long i = tid; i < CD; i += stride) {
      double s1 = 0.0, s2 = 0.0;
      const double x0 = (double)x[i];
      for (long long t = 0; t < T; ++t) {
        const double v = (double)x[t * CD + i] - x0;  // shifted sums: no cancellation
        s1 += v;
        s2 += v * v;
      }
      const double mean_sh = s1 / (double)T;
      const double var = (s2 - (double)T * mean_sh * mean_sh) / (double)(T - 1);
      const double mean = mean_sh + x0;
      sm += mean;
      sm2 += mean * mean;
      sv += var;
    }
    atomicAdd(&sh[d], sm);
    atomicAdd(&sh[D + d], sm2);
    atomicAdd(&sh[2 * D + d], sv);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) atomicAdd(&stats[i], sh[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[3 * D], (double)C);
}

// ---- ESS partial: chain-summed biased autocovariance for lag < num_lags.
// One block = one dimension d and up to 32 chains whose centred series sit in shared memory with an odd,
// zero-padded stride TS.  A warp owns a block of 8 consecutive lags, its lanes own the 32 series: per 8
// time steps a lane loads a[0..7] = p[t..t+7] and b[0..14] = p[t+l..t+l+14] from ITS series (bank =
// (9 s + t) mod 32: conflict-free) and issues 64 FMAs -- 0.36 shared loads per FMA with every thread
// busy (the first version ran one thread per lag: 2 loads per FMA and 64 active threads at 64 lags).
// The zero padding makes every out-of-range product vanish, so there is no bounds logic in the loop.
template <typename R>
__global__ void __launch_bounds__(256) k_ess_partial(const R* __restrict__ x, long long T, long long C, int D,
                                                     int num_lags, int S, int TS, double* acov) {
  extern __shared__ float xs[];  // S * TS
  const int d = blockIdx.y;
  const long long c0 = (long long)blockIdx.x * S;
  const int ns = (int)min((long long)S, C - c0);
  const long long CD = C * D;
  for (long long i = threadIdx.x; i < (long long)S * TS; i += blockDim.x) xs[i] = 0.f;
  __syncthreads();
  for (long long i = threadIdx.x; i < (long long)ns * T; i += blockDim.x) {
    const int s = (int)(i % ns);
    const long long t = i / ns;
    // shifted by the series' first sample IN THE SOURCE PRECISION before the cast (float64 inputs, or means much
    // larger than the spread, would otherwise lose their low bits here); the per-series mean is removed below
    const long long sd = (c0 + s) * D + d;
    xs[(long long)s * TS + t] = (float)(x[t * CD + sd] - x[sd]);
  }
  __syncthreads();
  // centre each series on its own mean (diagnostics.py:122-124)
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  for (int s = warp; s < ns; s += nwarp) {
    double sum = 0.0;
    for (long long t = lane; t < T; t += 32) sum += (double)xs[(long long)s * TS + t];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float m = (float)(sum / (double)T);
    for (long long t = lane; t < T; t += 32) xs[(long long)s * TS + t] -= m;
  }
  __syncthreads();
  const bool live = lane < ns;                       // S may be smaller than a warp
  const float* p = xs + (long long)(live ? lane : 0) * TS;
  for (int l0 = 8 * warp; l0 < num_lags; l0 += 8 * nwarp) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const long long n = T - l0;  // products with t >= T - l0 are all zero
    for (long long t = 0; t < n; t += 8) {
      float a[8], bb[15];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = p[t + i];
#pragma unroll
      for (int i = 0; i < 15; ++i) bb[i] = p[t + l0 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], bb[i + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double tot = live ? (double)acc[j] : 0.0;
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0 && l0 + j < num_lags) atomicAdd(&acov[(long long)(l0 + j) * D + d], tot / (double)T);
    }
  }
}


// ---- streaming R-hat / ESS: the same sufficient statistics without the (T, C, D) sample tensor ---------------
// Samples arrive in blocks [Tb, C, D] (the sample buffer of one fused launch, reused).  Per series (chain, dim) the
// accumulator keeps the shift x0 (its first sample), S = sum y, Q = sum y^2 (y = x - x0, float64), the first K and
// the last K shifted samples; per (lag, dim) it keeps the chain-summed raw lagged products P_k = sum_c sum_t
// y_t y_{t+k} (float64).  diagnostics.py:122-131 centres every series on ITS OWN mean m_c, which is only known at
// the end; with A = sum_{t < T-k} y_t = S - (sum of the last k), B = sum_{t >= k} y_t = S - (sum of the first k):
//     sum_t (y_t - m)(y_{t+k} - m) = P_k - m (A + B) + (T - k) m^2,
// so the correction needs only S, the head and the tail of each chain (k_stream_partial).
struct StreamWs {
  float* x0;      // [C, D]
  double* S;      // [C, D]
  double* Q;      // [C, D]
  float* head;    // [K, C, D]
  float* ring;    // [K, C, D] last K shifted samples, chronological
  double* P;      // [K, D]
};

__host__ __device__ inline long long stream_align(long long x) { return (x + 255) / 256 * 256; }

inline long long stream_carve(unsigned char* base, long long C, int D, int K, StreamWs* w) {
  long long off = 0;
  const long long CD = C * D;
  auto take = [&](long long bytes) { unsigned char* p = base ? base + off : nullptr; off += stream_align(bytes); return p; };
  StreamWs t;
  t.x0 = (float*)take(CD * 4);
  t.S = (double*)take(CD * 8);
  t.Q = (double*)take(CD * 8);
  t.head = (float*)take((long long)K * CD * 4);
  t.ring = (float*)take((long long)K * CD * 4);
  t.P = (double*)take((long long)K * D * 8);
  if (w) *w = t;
  return off;
}

// one block = one dimension d and up to 32 chains; shared memory holds [ring tail (K) | block (Tb) | zero pad] per
// series with an odd stride (same conflict-free lane = series mapping as k_ess_partial)
template <typename R>
__global__ void __launch_bounds__(256) k_stream_update(const R* __restrict__ x, long long Tb, long long C, int D, int K,
                                                       long long T_prev, int TS, StreamWs w) {
  extern __shared__ float xs[];  // 32 * TS
  const int d = blockIdx.y;
  const long long c0 = (long long)blockIdx.x * 32;
  const int ns = (int)min(32LL, C - c0);
  const long long CD = C * D;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  for (long long i = threadIdx.x; i < 32LL * TS; i += blockDim.x) xs[i] = 0.f;
  __syncthreads();
  // shifts: the first sample of the series
  for (int s = threadIdx.x; s < ns; s += blockDim.x) {
    const long long sd = (c0 + s) * D + d;
    if (T_prev == 0) w.x0[sd] = (float)x[sd];
  }
  __syncthreads();
  for (long long i = threadIdx.x; i < (long long)ns * K; i += blockDim.x) {
    const int s = (int)(i % ns);
    const long long k = i / ns;
    xs[(long long)s * TS + k] = T_prev == 0 ? 0.f : w.ring[k * CD + (c0 + s) * D + d];
  }
  for (long long i = threadIdx.x; i < (long long)ns * Tb; i += blockDim.x) {
    const int s = (int)(i % ns);
    const long long t = i / ns;
    const long long sd = (c0 + s) * D + d;
    const float y = (float)((double)x[t * CD + sd] - (double)w.x0[sd]);
    xs[(long long)s * TS + K + t] = y;
    if (T_prev + t < K) w.head[(T_prev + t) * CD + sd] = y;
  }
  __syncthreads();
  // S, Q (float64), warp per series
  for (int s = warp; s < ns; s += nwarp) {
    double s1 = 0.0, s2 = 0.0;
    for (long long t = lane; t < Tb; t += 32) {
      const double y = (double)xs[(long long)s * TS + K + t];
      s1 += y;
      s2 += y * y;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      const long long sd = (c0 + s) * D + d;
      w.S[sd] = (T_prev == 0 ? 0.0 : w.S[sd]) + s1;
      w.Q[sd] = (T_prev == 0 ? 0.0 : w.Q[sd]) + s2;
    }
  }
  // lagged products with the newer factor inside this block: sum_u y_u y_{u - l}, u in [K, K + Tb)
  const bool live = lane < ns;
  const float* p = xs + (long long)(live ? lane : 0) * TS;
  for (int l0 = 8 * warp; l0 < K; l0 += 8 * nwarp) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long u = K; u < K + Tb; u += 8) {
      float a[8], bb[15];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = p[u + i];
#pragma unroll
      for (int i = 0; i < 15; ++i) bb[i] = p[u - l0 - 7 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], bb[i - j + 7], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double tot = live ? (double)acc[j] : 0.0;
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (lane == 0 && l0 + j < K) atomicAdd(&w.P[(long long)(l0 + j) * D + d], tot);
    }
  }
  __syncthreads();
  // the last K samples become the tail of the next block
  for (long long i = threadIdx.x; i < (long long)ns * K; i += blockDim.x) {
    const int s = (int)(i % ns);
    const long long k = i / ns;
    w.ring[k * CD + (c0 + s) * D + d] = xs[(long long)s * TS + Tb + k];
  }
}

// chain-summed statistics in the layouts of k_rhat_partial (stats[3 D + 1]) and k_ess_partial (acov[num_lags, D]);
// block = one dimension d x 256 chains
__global__ void __launch_bounds__(256) k_stream_partial(long long T, long long C, int D, int K, int num_lags, StreamWs w,
                                                        double* stats, double* acov) {
  __shared__ double red[8];
  const int d = blockIdx.y, lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool live = c < C;
  const long long CD = C * D, sd = (live ? c : 0) * D + d;
  const double Td = (double)T;
  const double S = live ? w.S[sd] : 0.0, Q = live ? w.Q[sd] : 0.0;
  const double m = S / Td;
  auto block_sum = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
  };
  {
    const double mean = m + (live ? (double)w.x0[sd] : 0.0);
    const double var = (Q - Td * m * m) / (Td - 1.0);
    const double a = block_sum(live ? mean : 0.0), b = block_sum(live ? mean * mean : 0.0), v = block_sum(live ? var : 0.0);
    if (threadIdx.x == 0) {
      atomicAdd(&stats[d], a);
      atomicAdd(&stats[D + d], b);
      atomicAdd(&stats[2 * D + d], v);
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&stats[3 * D], (double)C);
  if (acov == nullptr) return;
  double headsum = 0.0, tailsum = 0.0;
  for (int k = 0; k < num_lags; ++k) {
    // head_k = first k samples, tail_k = last k samples
    const double corr = live ? (-m * (2.0 * S - tailsum - headsum) + (Td - (double)k) * m * m) : 0.0;
    const double tot = block_sum(corr);
    if (threadIdx.x == 0) atomicAdd(&acov[(long long)k * D + d], (tot + (blockIdx.x == 0 ? w.P[(long long)k * D + d] : 0.0)) / Td);
    if (live && k < K) {
      headsum += (double)w.head[(long long)k * CD + sd];
      tailsum += (double)w.ring[(long long)(K - 1 - k) * CD + sd];
    }
  }
}

}  // namespace gb

using namespace gb;

extern "C" {

int gb200_rhat_partial(const void* samples, int64_t T, int64_t C, int32_t D, double* stats, int32_t dtype, void* stream) {
  if (!samples || !stats || T < 2 || C < 1 || D < 1) { set_error("rhat_partial: bad argument (need T >= 2)"); return GB200_ERR_INVALID_ARGUMENT; }
  if (dtype != GB200_F32 && dtype != GB200_F64) { set_error("rhat_partial: bad dtype"); return GB200_ERR_INVALID_ARGUMENT; }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(stats, 0, sizeof(double) * (3 * D + 1), s);
  const int block = 256;
  long long want = (C * D + block - 1) / block;
  int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  if ((long long)grid * block < D) grid = (D + block - 1) / block;
  const size_t sh = sizeof(double) * 3 * D;
  if (dtype == GB200_F32) k_rhat_partial<float><<<grid, block, sh, s>>>((const float*)samples, T, C, D, stats);
  else k_rhat_partial<double><<<grid, block, sh, s>>>((const double*)samples, T, C, D, stats);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

int gb200_rhat_finalize(const double* st, int64_t T, int32_t D, double* rhat) {
  if (!st || !rhat || T < 2 || D < 1) { set_error("rhat_finalize: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  const double C = st[3 * D];
  if (C < 2) { set_error("potential_scale_reduction as implemented only works for two or more chains."); return GB200_ERR_INVALID_ARGUMENT; }
  for (int d = 0; d < D; ++d) {
    const double m = st[d] / C;
    const double var_means = (st[D + d] - C * m * m) / (C - 1.0);
    const double B = (double)T * var_means;
    const double W = st[2 * D + d] / C;
    rhat[d] = sqrt((B / W + (double)T - 1.0) / (double)T);
  }
  return GB200_OK;
}

int gb200_ess_partial(const void* samples, int64_t T, int64_t C, int32_t D, int32_t num_lags, double* acov,
                      int32_t dtype, void* stream) {
  if (!samples || !acov || T < 2 || C < 1 || D < 1 || num_lags < 1) { set_error("ess_partial: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  if (dtype != GB200_F32 && dtype != GB200_F64) { set_error("ess_partial: bad dtype"); return GB200_ERR_INVALID_ARGUMENT; }
  cudaStream_t s = (cudaStream_t)stream;
  if (num_lags > T) num_lags = (int32_t)T;
  cudaMemsetAsync(acov, 0, sizeof(double) * (size_t)num_lags * D, s);
  const size_t max_sh = 200 * 1024;
  const long long TS = ((T + 24) | 1);  // odd stride, >= 24 floats of zero padding behind every series
  long long S = (long long)(max_sh / (sizeof(float) * (size_t)TS));
  if (S < 1) { set_error("ess_partial: T=%lld too long for the shared-memory series tile", (long long)T); return GB200_ERR_UNSUPPORTED; }
  if (S > 32) S = 32;
  if (S > C) S = C;
  const size_t sh = sizeof(float) * (size_t)S * (size_t)TS;
  dim3 grid((unsigned)((C + S - 1) / S), (unsigned)D);
  const int block = 256;
  cudaError_t e;
  if (dtype == GB200_F32) {
    e = cudaFuncSetAttribute(k_ess_partial<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e == cudaSuccess) k_ess_partial<float><<<grid, block, sh, s>>>((const float*)samples, T, C, D, num_lags, (int)S, (int)TS, acov);
  } else {
    e = cudaFuncSetAttribute(k_ess_partial<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e == cudaSuccess) k_ess_partial<double><<<grid, block, sh, s>>>((const double*)samples, T, C, D, num_lags, (int)S, (int)TS, acov);
  }
  if (e != cudaSuccess) { set_error("ess_partial: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  GB_CHECK_LAUNCH();
  return GB200_OK;
}


// ---- streaming diagnostics (see StreamWs): N2 of SURVEY 8(f) ---------------------------------------------------
int64_t gb200_stream_diag_workspace(int64_t C, int32_t D, int32_t max_lags) {
  if (C < 1 || D < 1 || max_lags < 8) return -1;
  return stream_carve(nullptr, C, D, (max_lags + 7) / 8 * 8, nullptr) + 256;
}

int gb200_stream_diag_update(void* workspace, const void* samples, int64_t Tb, int64_t C, int32_t D, int32_t max_lags,
                             int64_t T_prev, int32_t dtype, void* stream) {
  if (!workspace || !samples || Tb < 1 || C < 1 || D < 1 || max_lags < 8 || T_prev < 0 || ((uintptr_t)workspace & 255)) {
    set_error("stream_diag_update: bad argument (workspace must be 256-byte aligned, max_lags >= 8)");
    return GB200_ERR_INVALID_ARGUMENT;
  }
  if (dtype != GB200_F32 && dtype != GB200_F64) { set_error("stream_diag_update: bad dtype"); return GB200_ERR_INVALID_ARGUMENT; }
  const int K = (max_lags + 7) / 8 * 8;
  StreamWs w;
  stream_carve((unsigned char*)workspace, C, D, K, &w);
  cudaStream_t s = (cudaStream_t)stream;
  if (T_prev == 0) cudaMemsetAsync(w.P, 0, sizeof(double) * (size_t)K * D, s);
  const long long TS = ((K + Tb + 24) | 1);
  const size_t sh = sizeof(float) * 32 * (size_t)TS;
  if (sh > 200 * 1024) { set_error("stream_diag_update: block of %lld samples + %d lags does not fit shared memory; use smaller blocks", (long long)Tb, K); return GB200_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)D);
  cudaError_t e;
  if (dtype == GB200_F32) {
    e = cudaFuncSetAttribute(k_stream_update<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e == cudaSuccess) k_stream_update<float><<<grid, 256, sh, s>>>((const float*)samples, Tb, C, D, K, T_prev, (int)TS, w);
  } else {
    e = cudaFuncSetAttribute(k_stream_update<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e == cudaSuccess) k_stream_update<double><<<grid, 256, sh, s>>>((const double*)samples, Tb, C, D, K, T_prev, (int)TS, w);
  }
  if (e != cudaSuccess) { set_error("stream_diag_update: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

int gb200_stream_diag_partial(void* workspace, int64_t T, int64_t C, int32_t D, int32_t max_lags, int32_t num_lags,
                              double* stats, double* acov, void* stream) {
  if (!workspace || !stats || T < 2 || C < 1 || D < 1 || max_lags < 8 || num_lags < 0 || ((uintptr_t)workspace & 255)) {
    set_error("stream_diag_partial: bad argument");
    return GB200_ERR_INVALID_ARGUMENT;
  }
  const int K = (max_lags + 7) / 8 * 8;
  if (num_lags > K) num_lags = K;
  if (num_lags > T) num_lags = (int32_t)T;
  StreamWs w;
  stream_carve((unsigned char*)workspace, C, D, K, &w);
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(stats, 0, sizeof(double) * (3 * D + 1), s);
  if (acov && num_lags > 0) cudaMemsetAsync(acov, 0, sizeof(double) * (size_t)num_lags * D, s);
  dim3 grid((unsigned)((C + 255) / 256), (unsigned)D);
  k_stream_partial<<<grid, 256, 0, s>>>(T, C, D, K, num_lags, w, stats, (acov && num_lags > 0) ? acov : nullptr);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

// Geyer initial positive + monotone sequence on the chain-averaged autocovariance
// (diagnostics.py:133-209), restricted to the lags that were computed.
int gb200_ess_finalize(const double* acov, const double* st, int64_t T, int64_t Ctot, int32_t D, int32_t num_lags,
                       double* ess, uint8_t* truncated) {
  if (!acov || !st || !ess || T < 4 || Ctot < 2 || D < 1 || num_lags < 2) { set_error("ess_finalize: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  if (num_lags > T) num_lags = (int32_t)T;
  const double N = (double)T, M = (double)Ctot;
  const long long n_even = T - T % 2;
  long long L = num_lags < n_even ? num_lags : n_even;
  L -= L % 2;
  const long long P = L / 2;       // number of (even, odd) pairs available
  const long long P_full = n_even / 2;
  std::vector<double> even(P), odd(P);
  std::vector<char> mask(P);
  for (int d = 0; d < D; ++d) {
    const double m = st[d] / M;
    const double var_means = (st[D + d] - M * m * m) / (M - 1.0);
    const double a0 = acov[d] / M;
    const double mean_var0 = a0 * N / (N - 1.0);
    const double weighted = mean_var0 * (N - 1.0) / N + var_means;
    for (long long k = 0; k < L; ++k) {
      const double rho = k == 0 ? 1.0 : 1.0 - (mean_var0 - acov[k * D + d] / M) / weighted;
      if (k % 2 == 0) even[k / 2] = rho; else odd[k / 2] = rho;
    }
    bool carry = true;
    long long max_t = 0;
    for (long long t = 0; t < P; ++t) {
      carry = carry && (even[t] + odd[t] > 0.0);
      if (carry) max_t = t;
      mask[t] = carry;
    }
    const bool trunc = carry && P < P_full;  // still positive at the last computed pair
    if (truncated) truncated[d] = trunc;
    // JAX: gather clamps out-of-bounds, scatter drops them
    const bool in_bounds = (max_t + 1) < P_full;
    const long long idx = (max_t + 1) < P ? (max_t + 1) : (P - 1);
    const double even_at_idx_raw = even[idx];
    for (long long t = 0; t < P; ++t) if (!mask[t]) odd[t] = 0.0;
    for (long long t = 0; t < P; ++t) {
      bool me = mask[t];
      if (t == idx && in_bounds && (max_t + 1) < P) me = even_at_idx_raw > 0.0;
      if (!me) even[t] = 0.0;
    }
    double prev = even[0] + odd[0];
    for (long long t = 0; t < P; ++t) {
      const double s = even[t] + odd[t];
      if (s > prev) { even[t] = prev / 2.0; odd[t] = prev / 2.0; }
      else prev = s;
    }
    double sum = 0.0;
    for (long long t = 0; t < P; ++t) sum += even[t] + odd[t];
    double tau = -1.0 + 2.0 * sum - even[idx];
    const double ess_raw = M * N;
    const double floor_tau = 1.0 / log10(ess_raw);
    if (tau < floor_tau) tau = floor_tau;
    ess[d] = ess_raw / tau;
  }
  return GB200_OK;
}

// Algorithmic FP32 flops per chain per integrator step of the closed-form algorithm
// (FMA = 2; one per div/sqrt/exp/log).  Derivation in DESIGN.md "Roofline accounting".
double gb200_flops_per_chain_step(int32_t sampler, const gb200_target_desc* t) {
  if (!t) return 0.0;
  const double D = t->D;
  if (t->kind == GB200_TARGET_FUNNEL) {
    switch (sampler) {
      case GB200_LMCMONGE: return 59.0 * D + 60.0;
      case GB200_LMC: return 25.0 * D + 140.0;
      case GB200_RMHMC: return 0.0;  // depends on fixed-point iterations; reported per f-eval separately
    }
  }
  return 0.0;
}

// Algorithmic FP32 flops per chain per TRANSITION outside the integrator steps (amortised over L by
// the caller): the normal transform of D uniforms (u, u^2, log1p: 2 divisions + ~22, branch offset,
// degree-8 Horner, scale: ~46 per normal), the draw, both energies and the sampler's prologue.
// Integer threefry work (75 ops per block) is NOT counted.  Derivation in DESIGN.md section 4.
double gb200_flops_per_transition(int32_t sampler, const gb200_target_desc* t) {
  if (!t) return 0.0;
  const double D = t->D;
  if (t->kind == GB200_TARGET_FUNNEL) {
    switch (sampler) {
      // normals 46D; L + normalise 3D+3; two HVPs 10D+12; rank-one Cholesky draw 11D; two kinetic
      // energies 8D+20; gradient rescale + accept D+5
      case GB200_LMCMONGE: return 79.0 * D + 40.0;
      // normals 46D; arrow-factor draw 4D; two kinetic energies 8D+20; accept ~10
      case GB200_LMC: return 58.0 * D + 30.0;
      case GB200_RMHMC: return 58.0 * D + 30.0;
    }
  }
  return 0.0;
}

}  // extern "C"
