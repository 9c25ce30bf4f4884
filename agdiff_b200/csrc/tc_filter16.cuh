// Arguments and small helpers shared by the CFConv kernels of the fp16-split family: tc_filter16.cu (one launch per conv, two
// groups each owning a tile end to end; A/B and unfused path) and tc_cfconv.cu (both convs in one warp-specialised launch - the default).
#pragma once
#include "tc16_common.cuh"

namespace agd {

constexpr int LDS_W = 68;               // padded row stride (floats) of the 64-column filter half-tile awaiting aggregation

struct TcF16Args {        // tc_filter16_kernel (unfused path)
  const uint32_t* W1img;   // [hi | lo'] fp16 images of F1 (K=128): each (128/64) x F rows x 128 B
  const uint32_t* W2img;   // ... of F2 (K=F)
  const float *f1b, *f2b, *beta_ptr;
  const float* cw;         // [E] envelope * distance weight of this conv (edge_weight_kernel)
  const float* wsc;        // [0] = 1/scale(F1), [1] = 1/scale(F2)
  const int* n_rows_dev;
  const uint4* g2h;        // pre-split encoder state (tc_common.cuh: g2h_index)
  float* filt;             // [E][192]
  int col0;                // 0 (conv1) or 128 (conv2): column offset in filt
  int scaled;              // 1: lo' scaled by 2^S + scale-input-d, 0: unscaled lo (A/B switch AGD_F16_LOSHIFT=0)
  int* range_flag;
};

// softplus(y) - ln2 with the argument already in log2 units (y2 = y * log2 e): ln2 * (log2(1 + 2^y2) - 1).  No threshold
// branch (F.softplus switches to the identity above 20): for 24 < y2 < 128 the sum 1 + 2^y2 rounds to 2^y2 and lg2 returns y2
// itself (abs. error ~2e-7), exactly what the linear branch yields; beyond 2^128 the chain gives +inf, which the fp16 split
// turns into the range flag and the host into a re-run on the 3xTF32 kernels (ssp_fast there keeps the threshold).
__device__ __forceinline__ float ssp_log2(float y2) {
  float e, sp;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y2));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(sp) : "f"(1.0f + e));
  return fmaf(sp, LN2F, -LN2F);
}

}  // namespace agd
