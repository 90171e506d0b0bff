// Shared between fisher_tc.cu (the warp-specialised tcgen05 GEMMs) and rmhmc_lockstep.cu (the sampler built on them).
#pragma once
#include <string.h>
#include "launch.h"

namespace gb {

constexpr int FT_M = 128;      // pairs (metric GEMM) / data rows (quadratic-form GEMM) per CTA: MMA M
constexpr int FT_N = 256;      // chains per CTA (MMA N).  With N = 128 the MMA operand fetch alone (A 4 KB + B 4 KB per
                               // 64-cycle tf32 MMA) saturates the 128 B/clk of shared memory and leaves nothing for the
                               // operand producers (measured: tensor pipe 48-53 %); N = 256 needs 96 B/clk and builds
                               // each Khatri-Rao A stage once per 256 chains instead of once per 128.
constexpr int FT_KT = 16;      // K extent of one stage
constexpr int FT_KC = 8;       // stages per TMEM accumulation chunk: 128 K steps (see the epilogue note)
constexpr int FT_THREADS = 288;  // warps 0-7: operand producers + accumulator drainers (2 threads per row); warp 8: TMA / MMA issuer
constexpr int FT_XS = FT_KT + 4;  // padded row stride (floats) of the staged X tile: conflict-free float4 row reads
constexpr int FT_LBO = 128;                  // bytes
constexpr int FT_SBO = (FT_KT / 4) * 128;    // bytes
constexpr int FT_A_BYTES = FT_M * FT_KT * 4;  // one A tile (hi or lo)
constexpr int FT_B_BYTES = FT_N * FT_KT * 4;  // one B tile (hi or lo)
constexpr int FT_NSA = 4;      // A stages in flight (ring): the producers run up to 4 K tiles ahead of the tensor core
constexpr int FT_NSB = 3;      // B slots (one bulk-TMA copy each, issued by a dedicated lane)
constexpr int FT_NXB = 4;      // X tile buffers of the metric GEMM (bulk-copied 4 K tiles ahead)
constexpr int FT_DRAIN_LAG = 2;  // a finished TMEM chunk is drained after the producers queued this many tiles of the next
constexpr int FT_STAGE_REGION = FT_NSA * 2 * FT_A_BYTES + FT_NSB * 2 * FT_B_BYTES;  // A ring + B slots (hi + lo each)
constexpr int FT_RS = FT_N + 4;                                           // row stride (floats) of the fused epilogue's s / R tile
constexpr int FT_EPI_REGION = FT_M * FT_RS * 4;                           // sT / R [128 data rows][FT_RS]
constexpr int FT_REGION = ((FT_STAGE_REGION > FT_EPI_REGION ? FT_STAGE_REGION : FT_EPI_REGION) + 1023) / 1024 * 1024;

// phases of a chain in the lock-step sampler (rmhmc_lockstep.cu)
enum { LS_PH_FIRST0 = 0, LS_PH_FIRST = 1, LS_PH_ITER = 2, LS_PH_EXPL = 3, LS_PH_END = 4, LS_PH_DONE = 5 };

__device__ __forceinline__ void ft_split(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);            // TF32-exact
  lo = __uint_as_float(__float_as_uint(a - hi) & 0xFFFFE000u);       // TF32-truncated remainder
}

struct FtArgs {
  const float* Xtile;        // metric GEMM: X re-tiled per K tile (fisher_xtile_kernel)
  int N, D;
  const unsigned char* Wt;   // B operand tiles (TF32 hi / lo, UMMA canonical layout), [chain tile][K tile]
  long long C;               // chains (columns), used when n_active == NULL
  const int* n_active;       // device: number of live columns (lock-step sampler); CTAs beyond it exit at once
  float alpha;
  float* out;                // metric: G [C, D, D] (packed == 0) or packed pairs [C, P]; quadratic forms: h [C, ldh]
  int packed;                // 0: dense G [C, D, D]; > 0: packed pairs, value = row stride in floats (>= P)
  int ksplit;                // metric GEMM, packed output only: the K range (data rows) is cut into ksplit parts (grid z);
  long long split_stride;    // part z writes its partial sums to out + z * split_stride; the consumer adds them (0/1 = off)
  const float* Xt;           // quadratic forms: X^T [D, ldx]
  int ldx;
  const short2* pairs;
  int PS, ldh;
  // fused epilogue of the quadratic-form GEMM (EPI == 1)
  const float* sbuf;         // sT[n, c] = sigmoid(eta) of data row n, chain slot c; row stride lds (% 4 == 0)
  long long lds;
  const float* y;
  const unsigned char* slot_phase;
  float* parts;              // [m tiles][D][Ccap]
  long long Ccap;
};

// launchers (fisher_tc.cu); all asynchronous on s
int ft_launch_xtile(const float* Xt, int ldx, int N, int D, float* Xtile, cudaStream_t s);
int ft_launch_pairs(short2* pairs, int D, cudaStream_t s);
int ft_launch_metric_gemm(const FtArgs& a, long long ctiles, cudaStream_t s);            // vec(G) = Z^T W
int ft_launch_quad_gemm(const FtArgs& a, long long ctiles, int epi, cudaStream_t s);     // h = Z vecsym(A)
// packed symmetric matrices Ap[c, P] (off-diagonal entries already doubled) -> B operand tiles of the quadratic-form GEMM
int ft_launch_quad_b_packed(const float* Ap, int D, long long C, const int* n_active, unsigned char* Bt, long long ctiles,
                            cudaStream_t s);
int ft_set_attributes(int D);  // cudaFuncSetAttribute for both GEMM kernels (outside stream capture)
// Split of the metric GEMM's K range that minimises waves x (K tiles per CTA + fixed cost) on `sms` SMs (one CTA per
// SM).  c4's shape: 3 pair tiles x 64 chain tiles = 192 CTAs = 1.3 waves of 148, i.e. two full-length waves; cut in
// two it is three waves of half the length.  Only taken when the model gains > 5 % (each part costs a pass
// over the packed output).
inline int ft_pick_ksplit(long long ctas, int ktiles, int sms) {
  int best = 1;
  double t1 = 0, tb = 0;
  for (int S = 1; S <= 4; ++S) {
    const int kper = (ktiles + S - 1) / S;
    if (kper < 8 || (long long)(S - 1) * kper >= ktiles) break;  // every part keeps at least one K tile
    const double t = (double)((ctas * S + sms - 1) / sms) * (kper + 12);  // fixed cost of a CTA (prologue, last drain, epilogue) ~ 12 K tiles
    if (S == 1) t1 = tb = t;
    else if (t < 0.95 * t1 && t < tb) { tb = t; best = S; }
  }
  return best;
}
inline int ft_ps(int D) { const int P = D * (D + 1) / 2; return (P + FT_KT - 1) / FT_KT * FT_KT; }
inline size_t ft_metric_smem(int D) { return (size_t)FT_REGION + FT_NXB * (size_t)D * FT_XS * 4 + 1024; }
inline size_t ft_quad_smem(int D) { return (size_t)FT_REGION + (size_t)FT_M * (D | 1) * 4 + 1024; }

}  // namespace gb
