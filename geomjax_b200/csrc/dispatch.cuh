// Layout dispatch: (EPL, LPC) = registers-per-lane x lanes-per-chain instantiations.
#pragma once
#include "transition.cuh"

namespace gb {

// X(EPL, LPC): every layout the library instantiates.  D <= EPL * LPC.
#define GB_LAYOUTS_1(X) X(2, 1) X(4, 1) X(8, 1) X(12, 1) X(20, 1) X(32, 1)
#define GB_LAYOUTS_2(X) X(6, 2) X(10, 2) X(16, 2)
#define GB_LAYOUTS_4(X) X(5, 4) X(8, 4) X(13, 4) X(25, 4)
#define GB_LAYOUTS_8(X) X(13, 8) X(20, 8) X(32, 8)
#define GB_LAYOUTS_32(X) X(16, 32) X(32, 32)
#define GB_LAYOUTS(X) GB_LAYOUTS_1(X) GB_LAYOUTS_2(X) GB_LAYOUTS_4(X) GB_LAYOUTS_8(X) GB_LAYOUTS_32(X)

// Each *_launch.cu is compiled once per lanes-per-chain group (-DGB_LPC=n) so that nvcc runs
// in parallel; GB_MY_LAYOUTS is the subset this translation unit instantiates.
#define GB_CAT_(a, b) a##b
#define GB_CAT(a, b) GB_CAT_(a, b)
// Layouts that additionally get compile-time-D ("exact", D == EPL*LPC) instantiations.
#define GB_EXACT_1(X) X(2, 1) X(20, 1)
#define GB_EXACT_2(X) X(10, 2)
#define GB_EXACT_4(X) X(5, 4) X(25, 4)
#define GB_EXACT_8(X)
#define GB_EXACT_32(X)
#ifdef GB_LPC
#define GB_MY_LAYOUTS(X) GB_CAT(GB_LAYOUTS_, GB_LPC)(X)
#define GB_MY_EXACT(X) GB_CAT(GB_EXACT_, GB_LPC)(X)
#define GB_LPC_NAME(base) GB_CAT(GB_CAT(base, _lpc), GB_LPC)
#endif

struct LayoutChoice { int epl, lpc; };

// Smallest-waste layout for D with the requested lanes-per-chain (0 = heuristic).
inline bool choose_layout(int D, int lpc_req, long long C, LayoutChoice* out) {
  static const LayoutChoice all[] = {
#define GB_X(e, l) {e, l},
      GB_LAYOUTS(GB_X)
#undef GB_X
  };
  const int n = (int)(sizeof(all) / sizeof(all[0]));
  if (lpc_req == 0) {
    // heuristic: keep <= ~24 elements per lane (register arrays) and prefer few lanes per
    // chain (no shuffle / redundant scalar work) when there are enough chains to fill the GPU.
    // measured on B200 (tools/sweep.py, funnel D=20, 65,536 chains): 2 lanes/chain beats 1 and 4
    // for lmcmonge (46 vs 56 vs 55 us/transition) and ties with 1 for lmc.
    if (D <= 12) lpc_req = 1;
    else if (D <= 32) lpc_req = 2;
    else if (D <= 100) lpc_req = 4;
    else if (D <= 256) lpc_req = 8;
    else lpc_req = 32;
  }
  int best = -1;
  for (int i = 0; i < n; ++i) {
    if (all[i].lpc != lpc_req) continue;
    if (all[i].epl * all[i].lpc < D) continue;
    if (best < 0 || all[i].epl < all[best].epl) best = i;
  }
  if (best < 0) {  // fall back to any layout that fits, fewest slots
    for (int i = 0; i < n; ++i) {
      if (all[i].epl * all[i].lpc < D) continue;
      if (best < 0 || all[i].epl * all[i].lpc < all[best].epl * all[best].lpc) best = i;
    }
  }
  if (best < 0) return false;
  *out = all[best];
  return true;
}

// grid/block for a sub-warp-per-chain kernel: small blocks so that the grid balances over
// 148 SMs (c2: 65,536 chains -> 2048 blocks of 32 threads = 13.8 blocks per SM).
inline void launch_shape(long long C, int lpc, int* grid, int* block) {
  const long long threads = C * lpc;
  int b = 128;
  while (b > 32 && (threads + b - 1) / b < 148LL * 16) b >>= 1;
  *block = b;
  *grid = (int)((threads + b - 1) / b);
}

}  // namespace gb
