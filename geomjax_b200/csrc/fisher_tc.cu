// Batched Fisher-information metrics over the chain dimension on the 5th-generation tensor cores:
//     G_c = X^T diag(w_c) X + alpha I,   w_c = s(1-s), s = sigmoid(X theta_c)        
// for ALL chains at once as ONE GEMM   vec(G)[P, C] = Z^T[P, N] . W[N, C],   Z[n,(i,j)] = x_ni x_nj,
// P = D(D+1)/2 pairs i <= j.  This is `metric_fn` of the logistic-regression target under vmap
// (what rmhmc/metrics.py:46,62,121 call at every evaluation).
//
// tcgen05 mapping (one CTA = 128 pairs x 128 chains, accumulators in TMEM):
//   A = Z^T tile  [M = 128 pairs ,  K = data rows]  K-major, built ON THE FLY in shared memory from a
//                 staged X tile (Z is never materialised: it would be N x P floats),
//   B = W^T tile  [N = 128 chains,  K = data rows]  K-major,
//   D = 128 lanes x 128 columns of FP32 in tensor memory, read back with tcgen05.ld.
// Precision: tcgen05 has no FP32 MMA.  Each operand is split into a TF32-exact high part and the
// TF32-truncated remainder; three MMAs (hi*hi + hi*lo + lo*hi) give ~2^-21 relative error per
// product ("3xTF32"), which keeps the metric within the 1e-5 parity budget.
// Shared-memory operand layout = UMMA canonical K-major, no swizzle: 8-row x 16-byte core matrices,
// [row group][k core][8 rows][16 B]: LBO (next core along K) = 128 B, SBO (next 8-row group) = KT/4*128 B.
#include "launch.h"

namespace gb {

constexpr int FT_M = 128;      // pairs per CTA (MMA M)
constexpr int FT_N = 128;      // chains per CTA (MMA N)
constexpr int FT_KT = 32;      // data rows per stage
constexpr int FT_KC = 4;       // stages per TMEM accumulation chunk (see the epilogue note)
constexpr int FT_THREADS = 128;
constexpr int FT_LBO = 128;                  // bytes
constexpr int FT_SBO = (FT_KT / 4) * 128;    // bytes
constexpr int FT_TILE_BYTES = FT_M * FT_KT * 4;

__device__ __forceinline__ uint64_t ft_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);             // start address, 16-byte units
  d |= (uint64_t)((FT_LBO >> 4) & 0x3FFF) << 16;      // leading byte offset (K direction)
  d |= (uint64_t)((FT_SBO >> 4) & 0x3FFF) << 32;      // stride byte offset (8-row groups)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  return d;                                           // layout_type = 0 (no swizzle), base_offset = 0
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t FT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FT_N >> 3) << 17) | ((uint32_t)(FT_M >> 4) << 24);

__device__ __forceinline__ void ft_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(FT_IDESC), "r"(accumulate)
      : "memory");
}

// byte offset of element (row, k) inside a canonical K-major tile
__device__ __forceinline__ int ft_off(int row, int k) {
  return (row >> 3) * FT_SBO + (k >> 2) * FT_LBO + (row & 7) * 16 + (k & 3) * 4;
}

__device__ __forceinline__ void ft_split(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);            // TF32-exact
  lo = __uint_as_float(__float_as_uint(a - hi) & 0xFFFFE000u);       // TF32-truncated remainder
}

// W[c, n] = s(1-s),  s = sigmoid(x_n . theta_c);  one thread per (chain, data row)
__global__ void fisher_weights_kernel(const float* __restrict__ Xt, int ldx, int N, int D, const float* __restrict__ theta,
                                      long long C, float* __restrict__ W, int ldw) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long c = idx / ldw;
  const int n = (int)(idx - c * ldw);
  if (c >= C) return;
  float w = 0.f;
  if (n < N) {
    float eta = 0.f;
    for (int i = 0; i < D; ++i) eta = fmaf(Xt[(size_t)i * ldx + n], theta[c * D + i], eta);
    const float s = 1.f / (1.f + expf(-eta));
    w = s * (1.f - s);
  }
  W[c * ldw + n] = w;  // rows N..ldw-1 are zero padding (K is processed in tiles of FT_KT)
}

__global__ void __launch_bounds__(FT_THREADS, 1)
fisher_metric_tc_kernel(const float* __restrict__ Xt, int ldx, int N, int D, const float* __restrict__ W, int ldw,
                        long long C, float alpha, float* __restrict__ G) {
  extern __shared__ __align__(1024) unsigned char ft_smem[];
  unsigned char* A_hi = ft_smem;
  unsigned char* A_lo = A_hi + FT_TILE_BYTES;
  unsigned char* B_hi = A_lo + FT_TILE_BYTES;
  unsigned char* B_lo = B_hi + FT_TILE_BYTES;
  float* xs = (float*)(B_lo + FT_TILE_BYTES);  // [D][FT_KT] staged X tile
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int P = D * (D + 1) / 2;
  const int m0 = blockIdx.x * FT_M;          // first pair of this CTA
  const long long c0 = (long long)blockIdx.y * FT_N;  // first chain of this CTA
  const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);

  // pair (i, j), i <= j, handled by this thread (row m0 + tid of the A tile)
  int pi = 0, pj = 0;
  {
    int m = m0 + tid;
    if (m < P) {
      int i = 0, rem = m;
      while (rem >= D - i) { rem -= D - i; ++i; }
      pi = i;
      pj = i + rem;
    } else {
      pi = -1;
    }
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "n"(FT_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(A_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(A_lo);
  const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(B_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(B_lo);
  // Two-level accumulation.  The tensor core adds partial products into TMEM with truncation, a
  // bias that grows linearly with the number of accumulated K steps (measured: 2.3e-5 relative at
  // N = 1000, 4.5e-5 at N = 2000).  Every FT_KC stages (128 data rows) the chunk is drained from TMEM
  // and added to FP32 register accumulators with round-to-nearest; the next chunk restarts at zero.
  float acc[FT_N];
#pragma unroll
  for (int e = 0; e < FT_N; ++e) acc[e] = 0.f;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t phase = 0;
  const int ktiles = (N + FT_KT - 1) / FT_KT;
  for (int kt = 0; kt < ktiles; ++kt) {
    const int n0 = kt * FT_KT;
    // stage the X tile [D][KT] (zero beyond N)
    for (int e = tid; e < D * FT_KT; e += FT_THREADS) {
      const int i = e / FT_KT, kk = e - i * FT_KT;
      xs[e] = (n0 + kk < N) ? Xt[(size_t)i * ldx + n0 + kk] : 0.f;
    }
    __syncthreads();
    // A tile: row = pair, z = x_i * x_j over the KT data rows
#pragma unroll 4
    for (int kk = 0; kk < FT_KT; ++kk) {
      const float a = (pi >= 0) ? xs[pi * FT_KT + kk] * xs[pj * FT_KT + kk] : 0.f;
      float hi, lo;
      ft_split(a, hi, lo);
      const int off = ft_off(tid, kk);
      *(float*)(A_hi + off) = hi;
      *(float*)(A_lo + off) = lo;
    }
    // B tile: row = chain, w over the KT data rows (W rows are padded with zeros up to ldw)
    {
      const long long c = c0 + tid;
      const float* wrow = W + (size_t)(c < C ? c : 0) * ldw + n0;
#pragma unroll 4
      for (int kk = 0; kk < FT_KT; ++kk) {
        const float b = (c < C && n0 + kk < ldw) ? wrow[kk] : 0.f;
        float hi, lo;
        ft_split(b, hi, lo);
        const int off = ft_off(tid, kk);
        *(float*)(B_hi + off) = hi;
        *(float*)(B_lo + off) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int k8 = 0; k8 < FT_KT / 8; ++k8) {
        const uint32_t adv = (uint32_t)k8 * 2u * FT_LBO;  // one MMA consumes 8 tf32 = 2 core matrices along K
        const uint32_t acc0 = ((kt % FT_KC) > 0 || k8 > 0) ? 1u : 0u;
        ft_mma(tmem_d, ft_smem_desc(a_hi_s + adv), ft_smem_desc(b_hi_s + adv), acc0);
        ft_mma(tmem_d, ft_smem_desc(a_hi_s + adv), ft_smem_desc(b_lo_s + adv), 1u);
        ft_mma(tmem_d, ft_smem_desc(a_lo_s + adv), ft_smem_desc(b_hi_s + adv), 1u);
      }
      // arrives on the mbarrier when every MMA issued so far has finished reading shared memory
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_a) : "memory");
    }
    // single-buffered: wait until the tensor core is done with this stage's operands
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(mbar_a), "r"(phase)
          : "memory");
    }
    phase ^= 1u;
    if ((kt % FT_KC) == FT_KC - 1 || kt == ktiles - 1) {
      // drain the chunk: TMEM -> registers (warp w owns lanes 32w..32w+31; thread = pair row)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int col = 0; col < FT_N; col += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem_d + lane_base + (uint32_t)col));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[col + e] += __uint_as_float(r[e]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // every warp has drained before the next chunk's first MMA overwrites TMEM
    }
  }

  // epilogue: registers -> G[c, i, j] (+ alpha on the diagonal)
  if (pi >= 0) {
#pragma unroll
    for (int e = 0; e < FT_N; ++e) {
      const long long c = c0 + e;
      if (c < C) {
        const float v = acc[e] + (pi == pj ? alpha : 0.f);
        float* g = G + (size_t)c * D * D;
        g[pi * D + pj] = v;
        g[pj * D + pi] = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(FT_N));
  }
}

}  // namespace gb

using namespace gb;

extern "C" int gb200_logreg_fisher_metric(const gb200_target_desc* t, const void* position, void* metric, void* workspace,
                                          int64_t workspace_bytes, int64_t C, int32_t dtype, void* stream) {
  if (!t || t->kind != GB200_TARGET_LOGREG) { set_error("fisher_metric: needs a logistic-regression target"); return GB200_ERR_INVALID_ARGUMENT; }
  if (dtype != GB200_F32) { set_error("fisher_metric: float32 only"); return GB200_ERR_UNSUPPORTED; }
  if (C == 0) return GB200_OK;
  if (!position || !metric || !workspace || C < 0 || !t->vec0) { set_error("fisher_metric: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  const int N = (int)t->N, D = t->D, ldx = (int)t->params[1];
  const int ldw = (N + FT_KT - 1) / FT_KT * FT_KT;
  const int64_t need = (int64_t)C * ldw * 4;
  if (workspace_bytes < need) { set_error("fisher_metric: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need); return GB200_ERR_INVALID_ARGUMENT; }
  cudaStream_t s = (cudaStream_t)stream;
  float* W = (float*)workspace;
  {
    const long long total = (long long)C * ldw;
    fisher_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const float*)t->vec0, ldx, N, D, (const float*)position, C, W, ldw);
    GB_CHECK_LAUNCH();
  }
  const int P = D * (D + 1) / 2;
  const size_t smem = 4 * FT_TILE_BYTES + (size_t)D * FT_KT * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(fisher_metric_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("fisher_metric: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  dim3 grid((unsigned)((P + FT_M - 1) / FT_M), (unsigned)((C + FT_N - 1) / FT_N));
  fisher_metric_tc_kernel<<<grid, FT_THREADS, smem, s>>>((const float*)t->vec0, ldx, N, D, W, ldw, C, (float)t->params[0], (float*)metric);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

extern "C" int64_t gb200_logreg_fisher_metric_workspace(const gb200_target_desc* t, int64_t C) {
  if (!t) return 0;
  const int ldw = ((int)t->N + FT_KT - 1) / FT_KT * FT_KT;
  return (int64_t)C * ldw * 4;
}
