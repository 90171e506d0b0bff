// Batched Fisher-information metrics over the chain dimension on the 5th-generation tensor cores:
//     G_c = X^T diag(w_c) X + alpha I,   w_c = s(1-s), s = sigmoid(X theta_c)        
// for ALL chains at once as ONE GEMM   vec(G)[P, C] = Z^T[P, N] . W[N, C],   Z[n,(i,j)] = x_ni x_nj,
// P = D(D+1)/2 pairs i <= j.  This is `metric_fn` of the logistic-regression target under vmap
// (what rmhmc/metrics.py:46,62,121 call at every evaluation).
//
// tcgen05 mapping (one CTA = 128 pairs x 256 chains, accumulators in TMEM):
//   A = Z^T tile  [M = 128 pairs ,  K = data rows]  K-major, built ON THE FLY in shared memory from a
//                 staged X tile (Z is never materialised: it would be N x P floats),
//   B = W^T tile  [N = 256 chains,  K = data rows]  K-major,
//   D = 128 lanes x 256 columns of FP32 in tensor memory, read back with tcgen05.ld.
// Precision: tcgen05 has no FP32 MMA.  Each operand is split into a TF32-exact high part and the
// TF32-truncated remainder; three MMAs (hi*hi + hi*lo + lo*hi) give ~2^-21 relative error per
// product ("3xTF32"), which keeps the metric within the 1e-5 parity budget.
// Shared-memory operand layout = UMMA canonical K-major, no swizzle: 8-row x 16-byte core matrices,
// [row group][k core][8 rows][16 B]: LBO (next core along K) = 128 B, SBO (next 8-row group) = KT/4*128 B.
#include "fisher_tc.cuh"

namespace gb {

__device__ __forceinline__ uint64_t ft_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);             // start address, 16-byte units
  d |= (uint64_t)((FT_LBO >> 4) & 0x3FFF) << 16;      // leading byte offset (K direction)
  d |= (uint64_t)((FT_SBO >> 4) & 0x3FFF) << 32;      // stride byte offset (8-row groups)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
  return d;                                           // layout_type = 0 (no swizzle), base_offset = 0
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 256
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

// ---- operand B: W^T tiles, pre-split and pre-laid-out --------------------------------------------------
// w[c, n] = s(1-s), s = sigmoid(x_n . theta_c), written ONCE per call as TF32 hi / lo parts directly in the
// UMMA canonical K-major tile layout, one 32 KB block per (chain tile of 256, K tile of 16 data rows):
// [hi tile 16 KB][lo tile 16 KB].  The GEMM kernel then fetches a B stage with ONE bulk-TMA copy instead
// of 128 threads issuing 32 scattered loads + splits each.  One thread = (chain, 4 consecutive data rows)
// = one 16-byte core-matrix row; 32 consecutive threads write one contiguous 512-byte row group.
__global__ void __launch_bounds__(256)
fisher_weights_kernel(const float* __restrict__ Xt, int ldx, int N, int D, const float* __restrict__ theta,
                      long long C, unsigned char* __restrict__ Wt, int ktiles, long long ctiles,
                      float* __restrict__ eta_out, long long ld_eta) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tile = gid >> 10;  // 1024 threads per (chain tile, K tile): 256 rows x 4 k-quads
  if (tile >= ctiles * ktiles) return;
  const int l = (int)(gid & 1023);
  const int rg = l >> 5, row = rg * 8 + (l & 7), kq = (l >> 3) & 3;  // a warp = one 8-row group = 512 contiguous bytes
  const long long ct = tile / ktiles;
  const int kt = (int)(tile - ct * ktiles);
  const long long c = ct * FT_N + row;
  const int n = kt * FT_KT + 4 * kq;
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C && n < N) {
    float eta[4] = {0.f, 0.f, 0.f, 0.f};
    const float* th = theta + c * D;
    for (int i = 0; i < D; ++i) {
      const float4 x = __ldg((const float4*)(Xt + (size_t)i * ldx + n));  // ldx % 4 == 0, columns >= N are zero
      const float t = __ldg(th + i);
      eta[0] = fmaf(x.x, t, eta[0]); eta[1] = fmaf(x.y, t, eta[1]);
      eta[2] = fmaf(x.z, t, eta[2]); eta[3] = fmaf(x.w, t, eta[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float sg = 1.f / (1.f + expf(-eta[e]));
      w[e] = (n + e < N) ? sg * (1.f - sg) : 0.f;
    }
    if (eta_out != nullptr) *(float4*)(eta_out + c * ld_eta + n) = make_float4(eta[0], eta[1], eta[2], eta[3]);
  }
  float4 hi, lo;
  ft_split(w[0], hi.x, lo.x); ft_split(w[1], hi.y, lo.y); ft_split(w[2], hi.z, lo.z); ft_split(w[3], hi.w, lo.w);
  unsigned char* base = Wt + (size_t)tile * (2 * FT_B_BYTES) + (size_t)rg * FT_SBO + kq * FT_LBO + (l & 7) * 16;
  *(float4*)base = hi;
  *(float4*)(base + FT_B_BYTES) = lo;
}

// X re-tiled per K tile: Xtile[kt][i][kk] = X[kt*32 + kk, i] (zero beyond N), so that the GEMM kernel fetches
// a whole [D][32] tile with one bulk copy (its per-tile staging loop was 30% of all stall samples).
__global__ void fisher_xtile_kernel(const float* __restrict__ Xt, int ldx, int N, int D, float* __restrict__ Xtile, int ktiles) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)ktiles * D * FT_XS) return;
  const int kk = (int)(e % FT_XS);
  const long long r = e / FT_XS;
  const int i = (int)(r % D), kt = (int)(r / D);
  const int n = kt * FT_KT + kk;
  Xtile[e] = (kk < FT_KT && n < N) ? Xt[(size_t)i * ldx + n] : 0.f;
}

__device__ __forceinline__ void ft_mbar_wait(uint32_t mbar_a, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar_a), "r"(parity)
        : "memory");
  }
}

// Fused epilogue of the quadratic-form GEMM (EPI == 1): out[i, c] = sum_r x[r, i] R[r, c] over the CTA's 128 data rows.
// Thread tile = TI features x TC chains (chains cb + (128 / TC) k: lanes read consecutive rows of Rs[c][129], no bank
// conflicts; the x reads are broadcasts), TI + TC shared-memory loads per TI * TC FMAs.
template <int TI, int TC>
__device__ __forceinline__ void ft_epi_xtr(const float* __restrict__ xr, int DX, const float* __restrict__ Rt, int D,
                                           float* __restrict__ parts, long long Ccap, long long c0, long long C, int tid) {
  constexpr int CB = 128 / TC;  // one call covers 128 chains; thread = TI features x TC chains (groups of 4 adjacent)
  const int cb = tid % CB, ib = tid / CB;
  if (ib * TI >= D) return;
  int fi[TI];
#pragma unroll
  for (int t = 0; t < TI; ++t) fi[t] = min(ib * TI + t, D - 1);
  float o[TI][TC];
#pragma unroll
  for (int t = 0; t < TI; ++t)
#pragma unroll
    for (int k = 0; k < TC; ++k) o[t][k] = 0.f;
  const float* rp = Rt + 4 * cb;  // chains 4 cb .. 4 cb + 3 (+ 4 CB per further group): float4 reads, lanes adjacent
#pragma unroll 4
  for (int r = 0; r < FT_M; ++r) {
    float xv[TI], rv[TC];
#pragma unroll
    for (int t = 0; t < TI; ++t) xv[t] = xr[r * DX + fi[t]];
#pragma unroll
    for (int k4 = 0; k4 < TC / 4; ++k4) {
      const float4 v = *(const float4*)(rp + r * FT_RS + k4 * 4 * CB);
      rv[4 * k4] = v.x; rv[4 * k4 + 1] = v.y; rv[4 * k4 + 2] = v.z; rv[4 * k4 + 3] = v.w;
    }
#pragma unroll
    for (int t = 0; t < TI; ++t)
#pragma unroll
      for (int k = 0; k < TC; ++k) o[t][k] = fmaf(xv[t], rv[k], o[t][k]);
  }
#pragma unroll
  for (int t = 0; t < TI; ++t) {
    const int i = ib * TI + t;
    if (i < D) {
#pragma unroll
      for (int k = 0; k < TC; ++k) {
        const long long j = c0 + (k >> 2) * 4 * CB + 4 * cb + (k & 3);
        if (j < C) parts[(size_t)i * Ccap + j] = o[t][k];
      }
    }
  }
}

// ---- the GEMM: warp-specialised, two A stages, three B slots, two TMEM accumulators -----------------------
// Producers (warps 0-7; thread pair = pair row = TMEM lane, each thread half of the K range and half of the
// accumulator columns), per K tile kt (16 data rows):
//   wait until the MMAs of tile kt-2 are done (A stage kt&1 and B slot (kt+1)%3 are free); thread 0 starts the
//   bulk copy of B(kt+1); wait for the X tile (bulk-copied two tiles ahead by the issuer); build the Khatri-Rao
//   A stage (z = x_i x_j, split into TF32 hi / lo) and ARRIVE on the stage's mbarrier -- no block-wide barrier.
//   On the first tile of a chunk they then drain the previous chunk's TMEM accumulator into FP32 registers
//   (two-level accumulation, see below) while the tensor core already works on the new chunk.
// Issuer (warp 8, one lane): waits for "A built" + "B landed", issues the 6 MMAs (3xTF32 x 2 K-steps, N = 256) from
//   pre-built descriptors, commits to the stage's "free" mbarrier (and the accumulator's "complete" mbarrier at
//   a chunk end), and refills the X buffer the producers just finished with.
// QUAD = false: vec(G)[pair, chain] = sum_n Z[n, pair] w[n, chain]   (M = pairs, K = data rows; A from X tiles)
// QUAD = true : h[n, chain] = sum_pair Z[n, pair] B[pair, chain]       (M = data rows, K = pairs; the 128 data rows
//               of the CTA are staged ONCE, the pair list comes from a table; B = m_pair * A_c[i, j] pre-tiled):
//               the quadratic forms x_n^T A_c x_n of rmhmc's dT/dq on the same pipeline.
// EPI (QUAD only): 0 = write h[c, n]; 1 = the lock-step sampler's fused epilogue (rmhmc_lockstep.cu):
//   R[n, c] = (slot_phase[c] == END ? 0 : 1/2 w'_n quad[n, c]) - (y_n - s[c, n]),  w' = s(1-s)(1-2s),
//   parts[m tile][i][c] = sum_{n in tile} x_ni R[n, c]   (summed over the m tiles by ls_reduce_kernel):
//   X^T R = dT/dq - X^T(y - s), the data part of dH/dq of rmhmc's implicit-midpoint map, without h or R ever
//   going to HBM.
#ifdef GB_FT_TIMING
__device__ unsigned long long ft_dbg[8192 * 16];
__device__ __forceinline__ unsigned long long ft_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define FT_STAMP(k) ft_dbg[(((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) % 4096 + (QUAD ? 4096 : 0)) * 16 + (k)] = ft_now()
#else
#define FT_STAMP(k)
#endif
template <bool QUAD, int EPI>
__global__ void __launch_bounds__(FT_THREADS, 1)
fisher_metric_tc_kernel(const FtArgs a) {
  if (threadIdx.x == 0) { FT_STAMP(0); }
  const int N = a.N, D = a.D;
  // split-K (metric GEMM, packed output): grid z = part of the K range; parts write partial sums side by side
  const int ktiles_all = QUAD ? a.PS / FT_KT : (N + FT_KT - 1) / FT_KT;
  const int kper = (!QUAD && a.ksplit > 1) ? (ktiles_all + a.ksplit - 1) / a.ksplit : ktiles_all;
  const int kt0 = QUAD ? 0 : (int)blockIdx.z * kper;
  const float* __restrict__ Xtile = a.Xtile + (size_t)kt0 * D * FT_XS;
  const unsigned char* __restrict__ Wt = a.Wt + (size_t)kt0 * (2 * FT_B_BYTES);
  const long long C = a.n_active ? (long long)*a.n_active : a.C;  // lock-step sampler: chains still iterating
  if ((long long)blockIdx.y * FT_N >= C) return;
  const float alpha = (QUAD || blockIdx.z == 0) ? a.alpha : 0.f;
  float* __restrict__ G = a.out + (QUAD ? 0 : (size_t)blockIdx.z * a.split_stride);
  const float* __restrict__ Xt = a.Xt;
  const int ldx = a.ldx;
  const short2* __restrict__ pairs = a.pairs;
  const int PS = a.PS, ldh = a.ldh;
  extern __shared__ __align__(1024) unsigned char ft_smem[];
  unsigned char* Astage = ft_smem;                                  // FT_NSA x [A_hi | A_lo]
  unsigned char* Bslot = ft_smem + FT_NSA * 2 * FT_A_BYTES;         // FT_NSB x [B_hi | B_lo] (one bulk copy each)
  float* xs0 = (float*)(ft_smem + FT_REGION);                       // FT_NXB x [D][FT_XS] X tiles / the staged data rows
  __shared__ uint32_t tmem_base_s;
  // mbarriers: A stage free (MMAs of its tile done) | A stage built | B slot landed | X tile landed | accumulator complete
  constexpr int MB_FREE = 0, MB_BUILT = MB_FREE + FT_NSA, MB_BLAND = MB_BUILT + FT_NSA, MB_XLAND = MB_BLAND + FT_NSB,
                MB_ACC = MB_XLAND + FT_NXB, MB_SLAND = MB_ACC + 2, MB_COUNT = MB_SLAND + 1;
  __shared__ __align__(8) unsigned long long mbar[MB_COUNT];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int P = D * (D + 1) / 2;
  const int m0 = blockIdx.x * FT_M;
  const long long ct = blockIdx.y;
  const long long c0 = ct * FT_N;
  const uint32_t mb0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]);
  auto mb = [&](int i) { return mb0 + 8u * (uint32_t)i; };
  const uint32_t xbytes = (uint32_t)D * FT_XS * 4;
  const uint32_t bbytes = 2 * FT_B_BYTES;
  const int ktiles = min(kper, ktiles_all - kt0);  // >= 1 (ft_pick_ksplit)
  const int nchunks = (ktiles + FT_KC - 1) / FT_KC;
  const int DX = D | 1;  // QUAD: odd row stride of the staged data rows xr[128][DX] (lanes = rows: conflict-free)
  if (QUAD) {
    // the CTA's 128 data rows, staged once, 16 loads in flight per thread
    float* xr = xs0;
    const int r = tid & (FT_M - 1), n = m0 + r;
    if (tid < 2 * FT_M) {
      const float* xp = Xt + min(n, N - 1);  // unconditional loads (clamped indices) so that all 16 are issued back to back
      for (int i0 = tid >> 7; i0 < D; i0 += 32) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldg(xp + (size_t)min(i0 + 2 * u, D - 1) * ldx);
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (i0 + 2 * u < D) xr[r * DX + i0 + 2 * u] = n < N ? v[u] : 0.f;
      }
    }
  }

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "n"(2 * FT_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < MB_COUNT; ++i) {
      const uint32_t cnt = (i >= MB_BUILT && i < MB_BLAND) ? 256u : (i == MB_SLAND ? 128u : 1u);  // "built": every producer thread arrives; "s landed": one per row
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb(i)), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  if (threadIdx.x == 0) { FT_STAMP(1); }

  auto bulk = [&](void* dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
  };

  if (warp == 8) {
    // ------------------------------------------------------------------ issuer (lane 0) and B loader (lane 1)
    if (tid == 257) {
      // B tile j -> slot j % FT_NSB as soon as the MMAs of tile j - FT_NSB (the slot's previous user) are done
      for (int j = 0; j < ktiles; ++j) {
        if (j >= FT_NSB) ft_mbar_wait(mb(MB_FREE + (j - FT_NSB) % FT_NSA), (uint32_t)(((j - FT_NSB) / FT_NSA) & 1));
        bulk(Bslot + (size_t)(j % FT_NSB) * bbytes, Wt + ((size_t)ct * ktiles_all + j) * bbytes, bbytes, mb(MB_BLAND + j % FT_NSB));
      }
    }
    if (tid == 256) {
      if (!QUAD) {
        for (int j = 0; j < FT_NXB && j < ktiles; ++j)
          bulk(xs0 + (size_t)j * D * FT_XS, Xtile + (size_t)j * D * FT_XS, xbytes, mb(MB_XLAND + j));
      }
      const uint64_t dA0 = ft_smem_desc((uint32_t)__cvta_generic_to_shared(Astage));
      const uint64_t dB0 = ft_smem_desc((uint32_t)__cvta_generic_to_shared(Bslot));
      constexpr uint64_t ATILE16 = FT_A_BYTES >> 4, BTILE16 = FT_B_BYTES >> 4;  // descriptor address units are 16 bytes
      for (int j = 0; j < ktiles; ++j) {
        const int s = j % FT_NSA, slot = j % FT_NSB;
        ft_mbar_wait(mb(MB_BUILT + s), (uint32_t)((j / FT_NSA) & 1));   // A stage built (and X buffer j % FT_NXB no longer read)
        if (!QUAD && j + FT_NXB < ktiles)
          bulk(xs0 + (size_t)(j % FT_NXB) * D * FT_XS, Xtile + (size_t)(j + FT_NXB) * D * FT_XS, xbytes, mb(MB_XLAND + j % FT_NXB));
        ft_mbar_wait(mb(MB_BLAND + slot), (uint32_t)((j / FT_NSB) & 1));  // B slot landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t a_hi = dA0 + (uint64_t)s * 2 * ATILE16, a_lo = a_hi + ATILE16;
        const uint64_t b_hi = dB0 + (uint64_t)slot * 2 * BTILE16, b_lo = b_hi + BTILE16;
        const int chunk = j / FT_KC;
        const uint32_t td = tmem_d + (uint32_t)((chunk & 1) * FT_N);
        if (j == 0) { FT_STAMP(2); }
        if (j == ktiles - 1) { FT_STAMP(3); }
#pragma unroll
        for (int k8 = 0; k8 < FT_KT / 8; ++k8) {
          const uint64_t adv = (uint64_t)k8 * ((2u * FT_LBO) >> 4);  // one MMA consumes 8 tf32 = 2 core matrices along K
          const uint32_t acc0 = ((j % FT_KC) > 0 || k8 > 0) ? 1u : 0u;
          ft_mma(td, a_hi + adv, b_hi + adv, acc0);
          ft_mma(td, a_hi + adv, b_lo + adv, 1u);
          ft_mma(td, a_lo + adv, b_hi + adv, 1u);
        }
        // arrive when every MMA issued so far has finished reading shared memory / writing TMEM
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb(MB_FREE + s)) : "memory");
        if ((j % FT_KC) == FT_KC - 1 || j == ktiles - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb(MB_ACC + (chunk & 1))) : "memory");
      }
    }
  } else {
    // ------------------------------------------------------------------ producers / drainers
    const int row = tid & 127, half = tid >> 7;
    if (QUAD && EPI == 1) {
      // The fused epilogue reads sT[row, chain] for this CTA's 128 rows x 256 chains (1 KB contiguous per row): ask
      // for those lines now, so that they wait in L2 when the main loop is done.  (Issued at kernel entry, these
      // prefetches queued in front of the data-row loads of the prologue.)
      for (int e = tid; e < FT_M * 8; e += 2 * FT_M) {
        const int n = m0 + (e >> 3);
        const long long j = c0 + (e & 7) * 32;
        if (j < a.lds && n < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.sbuf + (size_t)n * a.lds + j));
      }
    }
    int pi = 0, pj = 0;  // pair (i, j), i <= j, of row m0 + row
    if (!QUAD) {
      int m = m0 + row;
      if (m < P) {
        int i = 0, rem = m;
        while (rem >= D - i) { rem -= D - i; ++i; }
        pi = i;
        pj = i + rem;
      } else {
        pi = -1;
      }
    }
    // Two-level accumulation.  The tensor core adds partial products into TMEM with truncation, a
    // bias that grows linearly with the number of accumulated K steps (measured: 2.3e-5 relative at
    // N = 1000, 4.5e-5 at N = 2000).  Every FT_KC stages (128 K steps) the chunk is drained from TMEM
    // and added to FP32 register accumulators with round-to-nearest; the next chunk restarts at zero.
    constexpr int NH = FT_N / 2;  // accumulator columns (chains) owned by this thread
    float acc[NH];
#pragma unroll
    for (int e = 0; e < NH; ++e) acc[e] = 0.f;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;  // warp w may access TMEM lanes 32 (w % 4) ..

    auto drain = [&](int chunk) {  // accumulator (chunk & 1) -> registers
      ft_mbar_wait(mb(MB_ACC + (chunk & 1)), (uint32_t)((chunk >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t col0 = (uint32_t)((chunk & 1) * FT_N + half * NH);
#pragma unroll
      for (int col = 0; col < NH; col += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem_d + lane_base + col0 + (uint32_t)col));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[col + e] += __uint_as_float(r[e]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    int drained = 0;  // chunks already added to acc[]
    for (int kt = 0; kt < ktiles; ++kt) {
      const int s = kt % FT_NSA, use = kt / FT_NSA;
      const float* xs = xs0 + (size_t)(kt % FT_NXB) * D * FT_XS;
      unsigned char* A_hi = Astage + (size_t)s * 2 * FT_A_BYTES;
      unsigned char* A_lo = A_hi + FT_A_BYTES;
      if (kt >= FT_NSA) ft_mbar_wait(mb(MB_FREE + s), (uint32_t)((use - 1) & 1));  // MMAs of tile kt - FT_NSA done: A stage s is free
      if (QUAD) {
        // A stage: row = data row, z = x_i x_j over the 32 pairs of this K tile (pair list is warp-uniform)
        const float* xrow = xs0 + row * DX;
#pragma unroll
        for (int kq = 0; kq < FT_KT / 8; ++kq) {
          const int k4 = half * (FT_KT / 8) + kq;
          // four (i, j) pairs = one 16-byte load (warp-uniform address)
          const uint4 pq = __ldg((const uint4*)(pairs + kt * FT_KT + 4 * k4));
          const unsigned int pw[4] = {pq.x, pq.y, pq.z, pq.w};
          float z[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)  // padded pairs point at (0, 0) and meet zero rows of B
            z[e] = xrow[pw[e] & 0xffffu] * xrow[pw[e] >> 16];
          float4 hi, lo;
          ft_split(z[0], hi.x, lo.x); ft_split(z[1], hi.y, lo.y); ft_split(z[2], hi.z, lo.z); ft_split(z[3], hi.w, lo.w);
          const int off = (row >> 3) * FT_SBO + k4 * FT_LBO + (row & 7) * 16;
          *(float4*)(A_hi + off) = hi;
          *(float4*)(A_lo + off) = lo;
        }
      } else {
      ft_mbar_wait(mb(MB_XLAND + kt % FT_NXB), (uint32_t)((kt / FT_NXB) & 1));  // X tile kt landed
      {
        const float* xi = xs + (pi >= 0 ? pi : 0) * FT_XS;
        const float* xj = xs + (pi >= 0 ? pj : 0) * FT_XS;
        const bool ok = pi >= 0;
#pragma unroll
        for (int kq = 0; kq < FT_KT / 8; ++kq) {
          const int k4 = half * (FT_KT / 8) + kq;
          const float4 a4 = *(const float4*)(xi + 4 * k4);
          const float4 b4 = *(const float4*)(xj + 4 * k4);
          float4 hi, lo;
          ft_split(ok ? a4.x * b4.x : 0.f, hi.x, lo.x);
          ft_split(ok ? a4.y * b4.y : 0.f, hi.y, lo.y);
          ft_split(ok ? a4.z * b4.z : 0.f, hi.z, lo.z);
          ft_split(ok ? a4.w * b4.w : 0.f, hi.w, lo.w);
          const int off = (row >> 3) * FT_SBO + k4 * FT_LBO + (row & 7) * 16;
          *(float4*)(A_hi + off) = hi;
          *(float4*)(A_lo + off) = lo;
        }
      }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb(MB_BUILT + s)) : "memory");
      // FT_DRAIN_LAG tiles of a new chunk are queued: drain the previous chunk's accumulator under them.  This thread
      // arrives for a later tile only after its drain, and that accumulator's next overwrite (first tile of the chunk
      // after next, FT_KC > FT_DRAIN_LAG tiles later) is issued only after every producer arrived for that tile.
      if (kt >= FT_KC && (kt % FT_KC) == FT_DRAIN_LAG) {
        drain(kt / FT_KC - 1);
        drained = kt / FT_KC;
      }
    }
    if (tid == 0) { FT_STAMP(4); }
    for (int c = drained; c < nchunks; ++c) drain(c);
    if (tid == 0) { FT_STAMP(5); }

    if (QUAD && EPI == 1) {
      // Every MMA has completed (this thread drained the last accumulator): the operand stages are free.  The s tile
      // sT[128 rows][256 chains] is bulk-copied into them, one 1 KB row per thread, and R overwrites it in place.
      // (Read with per-thread global loads, s cost as much as the whole main loop at c4's shape: 268 us against
      // 131 us with the loads removed -- 8 KB in flight per SM cannot hide the latency; one bulk copy per row can.)
      float* St = (float*)ft_smem;  // [128 rows][FT_RS]: float4 accesses by lanes = rows and by lanes = chains are conflict-free
      const int n = m0 + row;
      if (half == 0) {
        const long long rem = a.lds - c0;  // lds % 4 == 0: whole 16-byte units
        const uint32_t nb = (n < N) ? (uint32_t)(rem < FT_N ? rem : FT_N) * 4u : 0u;
        if (nb) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb(MB_SLAND)), "r"(nb) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           (uint32_t)__cvta_generic_to_shared(St + row * FT_RS)),
                       "l"(a.sbuf + (size_t)n * a.lds + c0), "r"(nb), "r"(mb(MB_SLAND))
                       : "memory");
        } else {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb(MB_SLAND)) : "memory");
        }
      }
      const float yn = (n < N) ? __ldg(a.y + n) : 0.f;
      // phases of this thread's 128 chains: one byte per lane and a ballot per group of 32 (not 128 loads per thread)
      unsigned int endm[NH / 32];
#pragma unroll
      for (int gq = 0; gq < NH / 32; ++gq) {
        const long long jl = c0 + half * NH + gq * 32 + (tid & 31);
        endm[gq] = __ballot_sync(0xffffffffu, jl < C && __ldg(a.slot_phase + jl) == LS_PH_END);
      }
      if (tid == 0) { FT_STAMP(8); }
      ft_mbar_wait(mb(MB_SLAND), 0u);
      if (tid == 0) { FT_STAMP(9); }
      float* sp = St + row * FT_RS + half * NH;
#pragma unroll
      for (int e0 = 0; e0 < NH; e0 += 4) {  // fully unrolled: acc[] must keep static register indices
        const float4 s4 = *(const float4*)(sp + e0);
        const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
        float R[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long j = c0 + half * NH + e0 + u;
          R[u] = 0.f;
          if (j < C && n < N) {
            const float sg = sv[u];
            const float t = sg * (1.f - sg) * (1.f - 2.f * sg) * acc[e0 + u];
            R[u] = ((endm[(e0 + u) >> 5] >> ((e0 + u) & 31)) & 1u ? 0.f : 0.5f * t) - (yn - sg);
          }
        }
        *(float4*)(sp + e0) = make_float4(R[0], R[1], R[2], R[3]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 producer warps (the issuer warp is not involved)
      if (tid == 0) { FT_STAMP(10); }
      // parts[m tile][i][c] = sum_r x[r, i] R[r, c]: a 128 x D x 128 product on the FP32 pipe, register-tiled
      float* pbase = a.parts + (size_t)blockIdx.x * D * a.Ccap;
      if (D <= 32) {
        // both halves at once, 4 features x 8 chains per thread (threads 0-127: chains 0-127, 128-255: the rest):
        // the loop is bound by shared-memory wavefronts, 12 per 32 FMAs this way against 8 per 16 with a 4 x 4 tile
        const int h = tid >> 7;
        if (c0 + h * 128 < C) ft_epi_xtr<4, 8>(xs0, DX, St + h * 128, D, pbase, a.Ccap, c0 + h * 128, C, tid & 127);
      } else {
#pragma unroll 1
        for (int h = 0; h < FT_N / 128; ++h) {  // 128 chains at a time
          if (c0 + h * 128 >= C) break;
          ft_epi_xtr<8, 8>(xs0, DX, St + h * 128, D, pbase, a.Ccap, c0 + h * 128, C, tid);
        }
      }
    } else if (QUAD) {
      // epilogue: registers -> h[c, n] (chain-major rows of length ldh; lanes = consecutive data rows: coalesced)
      const int n = m0 + row;
      if (n < N) {
#pragma unroll
        for (int e = 0; e < NH; ++e) {
          const long long c = c0 + half * NH + e;
          if (c < C) G[(size_t)c * ldh + n] = acc[e];
        }
      }
    } else if (a.packed) {
      // epilogue, packed pairs G[c, m] (row stride a.packed, a multiple of 4 floats): the 128 x 256 tile goes through
      // the freed operand stages (every MMA has completed) and leaves as one bulk store of 512 bytes per chain.
      // (As 128 4-byte stores per thread this took 9-29 us per CTA at c4's shape, a third of the CTA's life.)
      float* Ss = (float*)ft_smem;  // [256 chains][128 pairs]
      const float dg = (pi >= 0 && pi == pj) ? alpha : 0.f;  // rows beyond P hold zeros (their A rows were zero)
#pragma unroll
      for (int e = 0; e < NH; ++e) Ss[(half * NH + e) * FT_M + row] = acc[e] + dg;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const long long c = c0 + tid;
      const int nrow = min(FT_M, a.packed - m0);
      if (c < C && nrow > 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(G + (size_t)c * a.packed + m0),
                     "r"((uint32_t)__cvta_generic_to_shared(Ss + tid * FT_M)), "r"((uint32_t)nrow * 4u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else if (pi >= 0) {
      // epilogue: registers -> G[c, i, j] (+ alpha on the diagonal)
#pragma unroll
      for (int e = 0; e < NH; ++e) {
        const long long c = c0 + half * NH + e;
        if (c < C) {
          const float v = acc[e] + (pi == pj ? alpha : 0.f);
          float* g = G + (size_t)c * D * D;
          g[pi * D + pj] = v;
          g[pj * D + pi] = v;
        }
      }
    }
  }
  if (tid == 0) { FT_STAMP(6); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) { FT_STAMP(7); }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * FT_N));
  }
}


// ---- operand preparation for the quadratic forms h[c, n] = x_n^T A_c x_n -------------------------------------
__global__ void quadform_pairs_kernel(short2* pairs, int D, int P, int PS) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= PS) return;
  int i = 0, rem = p;
  if (p < P) {
    while (rem >= D - i) { rem -= D - i; ++i; }
    pairs[p] = make_short2((short)i, (short)(i + rem));
  } else {
    pairs[p] = make_short2(0, 0);
  }
}
// B operand: row = chain, k = pair; value = A_c[i, j] (x 2 off the diagonal), TF32 hi / lo, UMMA tile layout
// (same thread -> address map as fisher_weights_kernel: 64 consecutive threads write one contiguous 1 KB group).
// PACKED: A is already the packed pair vector Ap[c, P] with the off-diagonal factor applied (ls_factor_kernel).
template <bool PACKED>
__global__ void __launch_bounds__(256)
quadform_b_kernel(const float* __restrict__ A, int D, int P, const short2* __restrict__ pairs, long long C_,
                  const int* __restrict__ n_active, unsigned char* __restrict__ Bt, int ktiles, long long ctiles) {
  const long long C = n_active ? (long long)*n_active : C_;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tile = gid >> 10;
  if (tile >= ctiles * ktiles) return;
  const int l = (int)(gid & 1023);
  const int rg = l >> 5, row = rg * 8 + (l & 7), kq = (l >> 3) & 3;  // a warp = one 8-row group = 512 contiguous bytes
  const long long ct = tile / ktiles;
  if (ct * FT_N >= C) return;
  const int kt = (int)(tile - ct * ktiles);
  const long long c = ct * FT_N + row;
  float w[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    if (PACKED) {
      const float* Ac = A + (size_t)c * P;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int p = kt * FT_KT + 4 * kq + e;
        if (p < P) w[e] = Ac[p];
      }
    } else {
      const float* Ac = A + (size_t)c * D * D;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int p = kt * FT_KT + 4 * kq + e;
        if (p < P) {
          const short2 ij = pairs[p];
          w[e] = Ac[ij.x * D + ij.y] * (ij.x == ij.y ? 1.f : 2.f);
        }
      }
    }
  }
  float4 hi, lo;
  ft_split(w[0], hi.x, lo.x); ft_split(w[1], hi.y, lo.y); ft_split(w[2], hi.z, lo.z); ft_split(w[3], hi.w, lo.w);
  unsigned char* base = Bt + (size_t)tile * (2 * FT_B_BYTES) + (size_t)rg * FT_SBO + kq * FT_LBO + (l & 7) * 16;
  *(float4*)base = hi;
  *(float4*)(base + FT_B_BYTES) = lo;
}

// ---- launchers ------------------------------------------------------------------------------------------
int ft_set_attributes(int D) {
  cudaError_t e = cudaFuncSetAttribute(fisher_metric_tc_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_metric_smem(D));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fisher_metric_tc_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_quad_smem(D));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fisher_metric_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ft_quad_smem(D));
  if (e != cudaSuccess) { set_error("tcgen05 GEMM: %s", cudaGetErrorString(e)); return GB200_ERR_CUDA; }
  return GB200_OK;
}
int ft_launch_xtile(const float* Xt, int ldx, int N, int D, float* Xtile, cudaStream_t s) {
  const int ktiles = (N + FT_KT - 1) / FT_KT;
  const long long total = (long long)ktiles * D * FT_XS;
  fisher_xtile_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Xt, ldx, N, D, Xtile, ktiles);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}
int ft_launch_pairs(short2* pairs, int D, cudaStream_t s) {
  const int P = D * (D + 1) / 2, PS = ft_ps(D);
  quadform_pairs_kernel<<<(PS + 255) / 256, 256, 0, s>>>(pairs, D, P, PS);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}
int ft_launch_metric_gemm(const FtArgs& a, long long ctiles, cudaStream_t s) {
  const int P = a.D * (a.D + 1) / 2;
  const int S = a.ksplit > 1 ? a.ksplit : 1;
  if (S > 1 && !a.packed) { set_error("metric GEMM: split-K needs the packed output"); return GB200_ERR_INVALID_ARGUMENT; }
  dim3 grid((unsigned)((P + FT_M - 1) / FT_M), (unsigned)ctiles, (unsigned)S);
  fisher_metric_tc_kernel<false, 0><<<grid, FT_THREADS, ft_metric_smem(a.D), s>>>(a);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}
int ft_launch_quad_gemm(const FtArgs& a, long long ctiles, int epi, cudaStream_t s) {
  if (ft_quad_smem(a.D) > 227 * 1024) { set_error("quadform: D=%d does not fit the staged data-row tile", a.D); return GB200_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((a.N + FT_M - 1) / FT_M), (unsigned)ctiles);
  if (epi) fisher_metric_tc_kernel<true, 1><<<grid, FT_THREADS, ft_quad_smem(a.D), s>>>(a);
  else fisher_metric_tc_kernel<true, 0><<<grid, FT_THREADS, ft_quad_smem(a.D), s>>>(a);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}
int ft_launch_quad_b_packed(const float* Ap, int D, long long C, const int* n_active, unsigned char* Bt, long long ctiles,
                            cudaStream_t s) {
  const int P = D * (D + 1) / 2, ktiles = ft_ps(D) / FT_KT;
  const long long total = ctiles * ktiles * 1024;
  quadform_b_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(Ap, D, P, nullptr, C, n_active, Bt, ktiles, ctiles);
  GB_CHECK_LAUNCH();
  return GB200_OK;
}

int fisher_metric_launch(const gb200_target_desc* t, const void* position, void* metric, void* workspace,
                         int64_t workspace_bytes, int64_t C, int32_t dtype, float* eta_out, long long ld_eta, void* stream);
}  // namespace gb

using namespace gb;

extern "C" int gb200_logreg_fisher_metric(const gb200_target_desc* t, const void* position, void* metric, void* workspace,
                                          int64_t workspace_bytes, int64_t C, int32_t dtype, void* stream) {
  return gb::fisher_metric_launch(t, position, metric, workspace, workspace_bytes, C, dtype, nullptr, 0, stream);
}
// eta_out (optional): eta[c, n] = x_n . theta_c, row stride ld_eta (multiple of 4, >= N rounded up to 4)
int gb::fisher_metric_launch(const gb200_target_desc* t, const void* position, void* metric, void* workspace,
                             int64_t workspace_bytes, int64_t C, int32_t dtype, float* eta_out, long long ld_eta, void* stream) {
  if (!t || t->kind != GB200_TARGET_LOGREG) { set_error("fisher_metric: needs a logistic-regression target"); return GB200_ERR_INVALID_ARGUMENT; }
  if (dtype != GB200_F32) { set_error("fisher_metric: float32 only"); return GB200_ERR_UNSUPPORTED; }
  if (C == 0) return GB200_OK;
  if (!position || !metric || !workspace || C < 0 || !t->vec0) { set_error("fisher_metric: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  const int N = (int)t->N, D = t->D, ldx = (int)t->params[1];
  const int ktiles = (N + FT_KT - 1) / FT_KT;
  const long long ctiles = (C + FT_N - 1) / FT_N;
  const int64_t wt_bytes = (int64_t)ctiles * ktiles * 2 * FT_B_BYTES;
  const int64_t need = wt_bytes + (int64_t)ktiles * D * FT_XS * 4;
  if (workspace_bytes < need) { set_error("fisher_metric: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need); return GB200_ERR_INVALID_ARGUMENT; }
  if (ldx % 4 != 0 || ((uintptr_t)workspace & 15) != 0) { set_error("fisher_metric: ldx must be a multiple of 4 and the workspace 16-byte aligned"); return GB200_ERR_INVALID_ARGUMENT; }
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* Wt = (unsigned char*)workspace;
  {
    const long long total = ctiles * ktiles * 1024;
    fisher_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const float*)t->vec0, ldx, N, D, (const float*)position, C, Wt, ktiles, ctiles, eta_out, ld_eta);
    GB_CHECK_LAUNCH();
  }
  float* Xtile = (float*)(Wt + wt_bytes);
  int rc = ft_launch_xtile((const float*)t->vec0, ldx, N, D, Xtile, s);
  if (rc) return rc;
  rc = ft_set_attributes(D);
  if (rc) return rc;
  FtArgs a;
  memset(&a, 0, sizeof(a));
  a.Xtile = Xtile; a.N = N; a.D = D; a.Wt = Wt; a.C = C; a.alpha = (float)t->params[0]; a.out = (float*)metric;
  return ft_launch_metric_gemm(a, ctiles, s);
}

// ---- quadratic forms h[c, n] = x_n^T A_c x_n for symmetric per-chain matrices A_c [C, D, D] ------------------
static int64_t quadform_ws(int D, int64_t C, int64_t* bt_bytes, int* PS_out) {
  const int PS = ft_ps(D);
  const int64_t ctiles = (C + FT_N - 1) / FT_N;
  const int64_t bt = ctiles * (PS / FT_KT) * 2 * FT_B_BYTES;
  if (bt_bytes) *bt_bytes = bt;
  if (PS_out) *PS_out = PS;
  return bt + (int64_t)PS * sizeof(short2);
}

extern "C" int64_t gb200_logreg_quadform_workspace(const gb200_target_desc* t, int64_t C) {
  return t ? quadform_ws(t->D, C, nullptr, nullptr) : 0;
}

extern "C" int gb200_logreg_quadform(const gb200_target_desc* t, const void* matrices, void* h, int64_t ldh, void* workspace,
                                     int64_t workspace_bytes, int64_t C, int32_t dtype, void* stream) {
  if (!t || t->kind != GB200_TARGET_LOGREG) { set_error("quadform: needs a logistic-regression target"); return GB200_ERR_INVALID_ARGUMENT; }
  if (dtype != GB200_F32) { set_error("quadform: float32 only"); return GB200_ERR_UNSUPPORTED; }
  if (C == 0) return GB200_OK;
  const int N = (int)t->N, D = t->D, ldx = (int)t->params[1];
  if (!matrices || !h || !workspace || C < 0 || !t->vec0 || ldh < N) { set_error("quadform: bad argument"); return GB200_ERR_INVALID_ARGUMENT; }
  int64_t bt_bytes;
  int PS;
  const int64_t need = quadform_ws(D, C, &bt_bytes, &PS);
  if (workspace_bytes < need || ((uintptr_t)workspace & 15) != 0) { set_error("quadform: workspace too small or not 16-byte aligned (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need); return GB200_ERR_INVALID_ARGUMENT; }
  const int P = D * (D + 1) / 2, ktiles = PS / FT_KT;
  const long long ctiles = (C + FT_N - 1) / FT_N;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* Bt = (unsigned char*)workspace;
  short2* pairs = (short2*)(Bt + bt_bytes);
  int rc = ft_launch_pairs(pairs, D, s);
  if (rc) return rc;
  {
    const long long total = ctiles * ktiles * 1024;
    quadform_b_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const float*)matrices, D, P, pairs, C, nullptr, Bt, ktiles, ctiles);
    GB_CHECK_LAUNCH();
  }
  rc = ft_set_attributes(D);
  if (rc) return rc;
  FtArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.D = D; a.Wt = Bt; a.C = C; a.out = (float*)h; a.Xt = (const float*)t->vec0; a.ldx = ldx; a.pairs = pairs; a.PS = PS;
  a.ldh = (int)ldh;
  return ft_launch_quad_gemm(a, ctiles, 0, s);
}

extern "C" int64_t gb200_logreg_fisher_metric_workspace(const gb200_target_desc* t, int64_t C) {
  if (!t) return 0;
  const int64_t ktiles = ((int64_t)t->N + FT_KT - 1) / FT_KT;
  const int64_t ctiles = (C + FT_N - 1) / FT_N;
  // pre-split W^T tiles (hi + lo), see fisher_weights_kernel, + the re-tiled X (fisher_xtile_kernel)
  return ctiles * ktiles * 2 * FT_B_BYTES + ktiles * (int64_t)t->D * FT_XS * 4;
}

#ifdef GB_FT_TIMING
extern "C" int gb200_debug_ft_stamps(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, gb::ft_dbg, sizeof(unsigned long long) * 8192 * 16);
}
#endif
