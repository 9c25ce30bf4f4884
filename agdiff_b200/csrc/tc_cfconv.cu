// Both CFConv layers of a SchNet interaction block (conv1: F = 128, conv2: F = 64; schnet.py:136-162,201-205) in ONE
// warp-specialised launch of the fp16-split tcgen05 family (numerics: header of tc_filter16.cu):
//
//   W1_e = ( F2a . SSP_b1( F1a . g2_e + b ) + b ) * cw1_e      agg[i][0:128]   = sum_{e -> i} x[src_e][0:128]   * W1_e
//   W2_e = ( F2b . SSP_b2( F1b . g2_e + b ) + b ) * cw2_e      agg[i][128:192] = sum_{e -> i} x[src_e][128:192] * W2_e
//
// Why one launch: both filter nets read the same encoder state g2_e (512 B per edge as the pre-split "g2h" tile), so the two
// separate launches of tc_filter16_kernel streamed it twice per block (12 x per evaluation) and paid the operand staging, the
// run bookkeeping and the tile barriers twice.  Why warp-specialised: in tc_filter16_kernel every 8-warp group walks a tile
// through all phases, so the warps that wait for an L2 gather cannot run an epilogue and no pipe gets above 40 % busy (ncu,
// profiles/r01b_ncu_full_f16_kernels.md).  Here every role has its own warps and they meet only through mbarriers:
//
//   warps  0- 7  E  epilogue 1 of both nets: accumulator -> bias + ShiftedSoftplus (SFU) -> fp16 hi/lo' split -> written back IN
//                   PLACE over the accumulator columns (a 32-column fp32 chunk becomes 16 hi + 16 lo' words), which is then the
//                   A operand of layer 2.  Nothing else: this is the MUFU-bound stage (2 per element).
//   warps  8-18  A  warps 8-15 drain the layer-2 accumulators (x bias, x edge weight) into a 64-column shared-memory half-tile;
//                   all eleven then run the CFConv aggregation: one warp per (destination run, 32-column slice), a quarter-warp
//                   per position-in-run mod 4, lanes = float4 columns - the summation order of cfconv_aggregate_kernel (four
//                   round-robin partial sums per destination, combined as (s0 + s1) + (s2 + s3)), which does not depend on where
//                   tile or CTA boundaries fall, so the result equals the unfused path bit for bit.  All tiles of a CTA pass
//                   through these warps in order; the partial sums of a run cut by a tile boundary wait in shared memory.
//   warp  19     M  weights (176 KB, cp.async.bulk, once) and every tcgen05.mma: executed warp-uniformly with elect.sync inside
//                   the instruction wrapper.
//   warps 20-23  L  operand loader: thread = tile row, g2h row (32 x LDG.128, L1 bypass) -> tcgen05.st into the layer-1 operand
//                   columns; the tile after next is pulled into L2 with cp.async.bulk.prefetch.
// Registers: launched at 80 per thread; the loader warpgroup gives registers back (setmaxnreg 48) and the three aggregation
// warpgroups take them (88): their gathers keep 36 rows x 16 B in flight per warp.
//
// TMEM (512 columns, one tile in flight, sub-tile pipelined):
//   [  0,128) A1   layer-1 operand hi | lo' (K = 128)          free again once layer 1 of both nets has completed
//   [128,256) X1   conv1: layer-1 accumulator, then (in place) layer-2 operand
//   [256,320) Y1   conv2: the same, 64 columns
//   [320,448) X2   conv1: layer-2 accumulator                   [448,512) Y2  conv2: layer-2 accumulator
// Tensor-pipe order per tile: L1x, L1y, (wait E) L2x, (wait E) L2y - while E works on X1 the pipe runs L1y, while it works on
// Y1 the pipe runs L2x; A drains tile j while the pipe and E are already on tile j+1.
#include "kernels.h"
#include "tc_filter16.cuh"

namespace agd {

using namespace tc;

constexpr int CF_THREADS = 768;                    // 24 warps x 80 registers
constexpr int CF_WARP_A = 8, CF_WARP_M = 19, CF_WARP_L = 20;
constexpr int CF_GROUP = 256;                      // E warps, and the draining A warps
constexpr int CF_AWARPS = CF_WARP_M - CF_WARP_A;   // 11 aggregation warps
constexpr int CF_ATHREADS = CF_AWARPS * 32;
constexpr uint32_t CFC_A1HI = 0, CFC_A1LO = 64, CFC_X1 = 128, CFC_Y1 = 256, CFC_X2 = 320, CFC_Y2 = 448;
constexpr uint32_t CF_W1X = 2u * 128u * 128u * 2u, CF_W1Y = 2u * 128u * 64u * 2u, CF_W2X = 2u * 128u * 128u * 2u, CF_W2Y = 2u * 64u * 64u * 2u;
constexpr int CF_STEPS = 9;                        // rows in flight per quarter-warp: 4 x 9 = 36 rows cover a radius-graph run (in-degree <= 33) in one round trip
enum { B_W = 0, B_A1_FULL, B_A1_FREE, B_D1X, B_D1Y, B_A2X, B_A2Y, B_D2X, B_D2Y, B_X2_FREE, B_Y2_FREE, B_COUNT };

struct CfArgs {
  const uint32_t *W1x, *W2x, *W1y, *W2y;   // [hi | lo'] fp16 operand images (pack.umma_image_f16)
  const float *b1x, *b2x, *b1y, *b2y;
  const float *beta_x, *beta_y;
  const float* wsc;                        // inverse weight scales [F1a, F2a, F1b, F2b]
  const float *cwx, *cwy;                  // [E] envelope * distance weight of conv1 / conv2 (edge_weight_kernel)
  const int* n_rows_dev;
  const uint4* g2h;
  int64_t g2h_blocks;                      // 128-row blocks backing g2h (prefetch bound)
  float* filt;                             // debug_filt: [E][192]
  int scaled;
  int* range_flag;
  int debug_filt;
  const float* xcat;
  float* agg;
  const int *e_src, *e_dst, *in_ptr;
};

constexpr size_t CF_SMEM = 1024 + CF_W1X + CF_W1Y + CF_W2X + CF_W2Y + (128 + 64 + 128 + 64) * sizeof(float) +
                           TM * LDS_W * sizeof(float) + 4 * 192 * sizeof(float) + (6 * TM + 8) * sizeof(int) + 16 * sizeof(uint64_t) + 64;

// ---- single-lane instructions issued from warp-uniform code (the elect lives inside the wrapper)
__device__ __forceinline__ void mma_f16_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
template <int SHIFT>
__device__ __forceinline__ void mma_f16_scaled_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p, %9;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(1u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "n"(SHIFT)
      : "memory");
}
__device__ __forceinline__ void mma_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One layer D[128 x N] = A[128 x K] . W^T as 3 x K/16 kind::f16 MMAs in the order of issue_3xf16 (cross terms first, folded
// into the main chain by scale-input-d).  INPLACE: the operand was written over a 32-column-chunked accumulator (epilogue 1):
// k-block kb has its hi words at a + 32 (kb / 2) + 8 (kb % 2) and its lo' words 16 columns further.
template <int K, int N, bool INPLACE>
__device__ __forceinline__ void cf_issue(uint32_t D, uint32_t a_hi, uint32_t a_lo, uint32_t w_smem, bool scaled) {
  constexpr uint32_t idesc = idesc_f16(N);
  constexpr uint32_t half_bytes = static_cast<uint32_t>(K) * N * 2u;
  const uint64_t d_hi = smem_desc_sw128(w_smem), d_lo = smem_desc_sw128(w_smem + half_bytes);
  auto hi_at = [&](int kb) { return INPLACE ? a_hi + 32u * (kb >> 1) + 8u * (kb & 1) : a_hi + 8u * kb; };
  auto lo_at = [&](int kb) { return INPLACE ? a_hi + 32u * (kb >> 1) + 8u * (kb & 1) + 16u : a_lo + 8u * kb; };
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    mma_f16_e(D, hi_at(kb), d_lo + boff16, idesc, kb > 0 ? 1u : 0u);
    mma_f16_e(D, lo_at(kb), d_hi + boff16, idesc, 1u);
  }
  if (scaled) mma_f16_scaled_e<F16_LO_SHIFT>(D, hi_at(0), d_hi, idesc);
  else mma_f16_e(D, hi_at(0), d_hi, idesc, 1u);
#pragma unroll
  for (int kb = 1; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    mma_f16_e(D, hi_at(kb), d_hi + boff16, idesc, 1u);
  }
}

// epilogue 1 of one 32-column accumulator chunk, in place: t = SSP(D / s1 + b1) -> 16 hi words | 16 lo' words
__device__ __forceinline__ void cf_epi1_chunk(uint32_t taddr, const float* s_b, float inv1, float lo_scale, __half2& amax) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  wait_ld();
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float t0 = ssp_log2(fmaf(__uint_as_float(v[2 * q]), inv1, s_b[2 * q]));
    const float t1 = ssp_log2(fmaf(__uint_as_float(v[2 * q + 1]), inv1, s_b[2 * q + 1]));
    split2_f16(t0, t1, lo_scale, hi[q], lo[q], amax);
  }
  tmem_st16(taddr, hi);
  tmem_st16(taddr + 16, lo);
}

__global__ void __launch_bounds__(CF_THREADS, 1) tc_cfconv_kernel(const CfArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w1x = base;
  uint8_t* w1y = w1x + CF_W1X;
  uint8_t* w2x = w1y + CF_W1Y;
  uint8_t* w2y = w2x + CF_W2X;
  float* s_b1x = reinterpret_cast<float*>(w2y + CF_W2Y);   // [128] pre-multiplied by beta * log2 e
  float* s_b1y = s_b1x + 128;                              // [64]
  float* s_b2x = s_b1y + 64;                               // [128]
  float* s_b2y = s_b2x + 128;                              // [64]
  float* s_W = s_b2y + 64;                                 // [128][LDS_W] filter half-tile awaiting aggregation
  float* s_carry = s_W + TM * LDS_W;                       // [4][192] partial sums of the run cut by the last tile boundary
  int* s_bk = reinterpret_cast<int*>(s_carry + 4 * 192);   // [2 tile parities][3][128] x-row offsets (src * 192) | destinations | run starts
  int* s_meta = s_bk + 2 * 3 * TM;                         // [2][4] runs in the tile, carry in, carry out, tile row where run 0 began
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_meta + 8);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = *a.n_rows_dev;
  // this CTA's contiguous row range, snapped to run (destination) boundaries - the partition of tc_filter16_kernel<F, true>
  const int64_t rows_per_cta = ((static_cast<int64_t>(n_rows) + TM - 1) / TM + gridDim.x - 1) / gridDim.x * TM;
  int64_t cta_begin = static_cast<int64_t>(blockIdx.x) * rows_per_cta, cta_end = cta_begin + rows_per_cta;
  if (cta_begin > n_rows) cta_begin = n_rows;
  if (cta_end > n_rows) cta_end = n_rows;
  if (cta_begin > 0 && cta_begin < n_rows) cta_begin = __ldg(a.in_ptr + __ldg(a.e_dst + cta_begin));
  if (cta_end < n_rows) cta_end = __ldg(a.in_ptr + __ldg(a.e_dst + cta_end));
  const int T = static_cast<int>((cta_end - cta_begin + TM - 1) / TM);   // tiles of this CTA (<= 0: nothing to do)

  if (warp == 0) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(&bars[B_W], 1);
    mbar_init(&bars[B_A1_FULL], 128);
    mbar_init(&bars[B_A1_FREE], 1);
    mbar_init(&bars[B_D1X], 1);
    mbar_init(&bars[B_D1Y], 1);
    mbar_init(&bars[B_A2X], CF_GROUP);
    mbar_init(&bars[B_A2Y], CF_GROUP);
    mbar_init(&bars[B_D2X], 1);
    mbar_init(&bars[B_D2Y], 1);
    mbar_init(&bars[B_X2_FREE], CF_GROUP);
    mbar_init(&bars[B_Y2_FREE], CF_GROUP);
    fence_barrier_init();
  }
  if (tid >= 64 && tid < 64 + 128) {
    const int i = tid - 64;
    s_b1x[i] = __ldg(a.b1x + i) * (__ldg(a.beta_x) * 1.4426950408889634f);
    s_b2x[i] = __ldg(a.b2x + i);
    if (i < 64) {
      s_b1y[i] = __ldg(a.b1y + i) * (__ldg(a.beta_y) * 1.4426950408889634f);
      s_b2y[i] = __ldg(a.b2y + i);
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const bool scaled = a.scaled != 0;
  // register rebalancing, one instruction per warpgroup: warpgroups 2-4 (aggregation warps 8-18 + the MMA warp) take what the
  // loader warpgroup (5) gives back; the epilogue warpgroups (0, 1) keep the launch allocation of 80
  if (warp >= 20) reg_dec<48>();
  else if (warp >= 8) reg_inc<88>();

  if (warp < CF_WARP_A) {
    // ================================================================== E: epilogue 1 of both nets
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float inv1x = __ldg(a.wsc + 0) * (__ldg(a.beta_x) * 1.4426950408889634f);
    const float inv1y = __ldg(a.wsc + 2) * (__ldg(a.beta_y) * 1.4426950408889634f);
    const float lo_scale = scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    for (int j = 0; j < T; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      mbar_wait(&bars[B_D1X], ph);
      fence_after_sync();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n0 = half * 64 + c * 32;
        cf_epi1_chunk(trow + CFC_X1 + n0, s_b1x + n0, inv1x, lo_scale, amax);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive(&bars[B_A2X]);
      mbar_wait(&bars[B_D1Y], ph);
      fence_after_sync();
      cf_epi1_chunk(trow + CFC_Y1 + half * 32, s_b1y + half * 32, inv1y, lo_scale, amax);
      wait_st();
      fence_before_sync();
      mbar_arrive(&bars[B_A2Y]);
    }
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
  } else if (warp < CF_WARP_M) {
    // ================================================================== A: drain of layer 2 + aggregation
    const int gwarp = warp - CF_WARP_A, gtid = tid - CF_WARP_A * 32;
    const bool drainer = gwarp < 8, helper = !drainer;
    const int quad = gwarp & 3, half = (gwarp >> 2) & 1;
    const int my_row = quad * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float inv2x = __ldg(a.wsc + 1), inv2y = __ldg(a.wsc + 3);
    const int rq = lane >> 3, c4 = (lane & 7) * 4;   // aggregation: position-in-run mod 4 and float4 column of this lane

    // Bookkeeping of tile jn -> s_bk[jn & 1] (helper warps 16-18, while the draining warps wait for the tensor pipe): x-row
    // offsets and destinations of the rows, run starts, and whether the first / last run continues across the tile boundary.
    auto bookkeep = [&](int jn) {
      int* bsrc = s_bk + (jn & 1) * 3 * TM;
      int* bdst = bsrc + TM;
      int* bruns = bdst + TM;
      const int64_t rown = cta_begin + static_cast<int64_t>(jn) * TM;
      const int nv = (cta_end - rown < TM) ? static_cast<int>(cta_end - rown) : TM;
      for (int i = gtid - 8 * 32; i < TM; i += 3 * 32) {
        bsrc[i] = (i < nv) ? __ldg(a.e_src + rown + i) * 192 : 0;
        bdst[i] = (i < nv) ? __ldg(a.e_dst + rown + i) : -1;
      }
      group_sync(2, 3 * 32);
      if (gwarp == 8) {
        const int prev_dst = (jn > 0) ? __ldg(a.e_dst + rown - 1) : -2;
        const int next_dst = (rown + nv < cta_end) ? __ldg(a.e_dst + rown + nv) : -5;
        int n_runs = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int row = w * 32 + lane;
          const int d = bdst[row];
          const int dp = (row > 0) ? bdst[row - 1] : -3;
          const bool start = row < nv && (row == 0 || d != dp);
          const uint32_t m = __ballot_sync(0xffffffffu, start);
          if (start) bruns[n_runs + __popc(m & ((1u << lane) - 1u))] = row;
          n_runs += __popc(m);
        }
        const int d0 = bdst[0];
        const bool cin = (d0 == prev_dst);
        if (lane == 0) {
          int* meta = s_meta + (jn & 1) * 4;
          meta[0] = n_runs;
          meta[1] = cin ? 1 : 0;
          meta[2] = (next_dst == bdst[nv - 1]) ? 1 : 0;
          // tile row at which run 0 began (<= 0: in an earlier tile): positions in a run are counted from its first edge
          meta[3] = cin ? static_cast<int>(static_cast<int64_t>(__ldg(a.in_ptr + d0)) - rown) : 0;
        }
      }
    };

    // the x rows of one aggregation item (run k, 32-column slice sl) that this lane multiplies: rows first, first + 4, ...
    struct Item { int e, first, col, colw; bool cin, cout, any; int dst_row; };
    float4 xv[CF_STEPS];
    auto item_of = [&](int item, int jt, int pass) {
      const int* bruns = s_bk + (jt & 1) * 3 * TM + 2 * TM;
      const int* meta = s_meta + (jt & 1) * 4;
      const int64_t rowt = cta_begin + static_cast<int64_t>(jt) * TM;
      const int nv = (cta_end - rowt < TM) ? static_cast<int>(cta_end - rowt) : TM;
      const int n_runs = meta[0];
      Item it;
      it.any = item < 2 * n_runs;
      const int k = item >> 1, sl = item & 1;
      const int s = it.any ? bruns[k] : 0;
      it.e = it.any ? ((k + 1 < n_runs) ? bruns[k + 1] : nv) : 0;
      const int base = (k == 0) ? meta[3] : s;
      it.first = s + ((rq - (s - base)) & 3);
      it.colw = sl * 32 + c4;
      it.col = ((pass == 2) ? 128 : pass * 64) + it.colw;
      it.cin = (k == 0) && meta[1] != 0;
      it.cout = (k == n_runs - 1) && meta[2] != 0;
      it.dst_row = s;
      return it;
    };
    auto load_x = [&](const Item& it, int row, int jt) {
      const int* bsrc = s_bk + (jt & 1) * 3 * TM;
#pragma unroll
      for (int u = 0; u < CF_STEPS; ++u) {
        const int rr = row + 4 * u;
        if (rr < it.e) xv[u] = __ldg(reinterpret_cast<const float4*>(a.xcat + bsrc[rr] + it.col));
      }
    };

    if (T > 0) {
      if (helper) bookkeep(0);
      group_sync(1, CF_ATHREADS);
      const Item it0 = item_of(gwarp, 0, 0);
      load_x(it0, it0.first, 0);
    }
    for (int j = 0; j < T; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      const int64_t r = row0 + my_row;
      const bool valid = my_row < n_valid;
      const int* bdst = s_bk + (j & 1) * 3 * TM + TM;
      const int n_items = 2 * s_meta[(j & 1) * 4];
      float cwx = 0.f, cwy = 0.f;
      if (drainer && valid) {
        cwx = __ldg(a.cwx + r);
        cwy = __ldg(a.cwy + r);
      }
      // three 64-column passes: conv1 columns [0,64), [64,128), conv2 columns [0,64) (= xcat / agg columns 128..191)
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const bool isy = pass == 2;
        if (helper && pass == 0 && j + 1 < T) bookkeep(j + 1);
        if (drainer) {
          const int n0 = (isy ? 0 : pass * 64) + half * 32;            // this thread's 32 filter columns of the net
          if (pass == 0) mbar_wait(&bars[B_D2X], ph);
          if (isy) mbar_wait(&bars[B_D2Y], ph);
          fence_after_sync();
          const float inv2 = isy ? inv2y : inv2x, cw = isy ? cwy : cwx;
          const float* sb = (isy ? s_b2y : s_b2x) + n0;
          float4* dstW = reinterpret_cast<float4*>(s_W + my_row * LDS_W + half * 32);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t v[16];
            tmem_ld16(trow + (isy ? CFC_Y2 : CFC_X2) + n0 + 16 * cc, v);
            wait_ld();
            if (cc == 1 && pass >= 1) {   // the accumulator of this net is drained: layer 2 of the next tile may overwrite it
              fence_before_sync();
              mbar_arrive(&bars[isy ? B_Y2_FREE : B_X2_FREE]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int n = 16 * cc + 4 * q;
              float4 o;
              o.x = fmaf(__uint_as_float(v[q * 4 + 0]), inv2, sb[n + 0]) * cw;
              o.y = fmaf(__uint_as_float(v[q * 4 + 1]), inv2, sb[n + 1]) * cw;
              o.z = fmaf(__uint_as_float(v[q * 4 + 2]), inv2, sb[n + 2]) * cw;
              o.w = fmaf(__uint_as_float(v[q * 4 + 3]), inv2, sb[n + 3]) * cw;
              dstW[4 * cc + q] = o;
              if (a.debug_filt && valid) *reinterpret_cast<float4*>(a.filt + r * 192 + (isy ? 128 : pass * 64) + half * 32 + n) = o;
            }
          }
        }
        group_sync(1, CF_ATHREADS);   // half-tile (and, before pass 0 ends, the next tile's bookkeeping) complete
        // items = (run, 32-column slice).  Quarter-warp rq takes the rows whose position in the run is rq mod 4, in order;
        // lanes = float4 columns.  Summation order == cfconv_aggregate_kernel (schnet.cu).  The x rows of this warp's first
        // item were requested before the previous barrier (below), so their L2 latency is hidden behind barrier + drain.
        for (int item = gwarp; item < n_items; item += CF_AWARPS) {
          const Item it = item_of(item, j, pass);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (it.cin) acc = *reinterpret_cast<const float4*>(s_carry + rq * 192 + it.col);
          for (int row = it.first; row < it.e; row += 4 * CF_STEPS) {
            if (item != gwarp || row != it.first) load_x(it, row, j);
#pragma unroll
            for (int u = 0; u < CF_STEPS; ++u) {
              const int rr = row + 4 * u;
              if (rr < it.e) {
                const float4 w = *reinterpret_cast<const float4*>(s_W + rr * LDS_W + it.colw);
                acc.x = fmaf(xv[u].x, w.x, acc.x);
                acc.y = fmaf(xv[u].y, w.y, acc.y);
                acc.z = fmaf(xv[u].z, w.z, acc.z);
                acc.w = fmaf(xv[u].w, w.w, acc.w);
              }
            }
          }
          if (it.cout) {
            *reinterpret_cast<float4*>(s_carry + rq * 192 + it.col) = acc;
          } else {   // (s0 + s1) + (s2 + s3): xor 8 pairs quarter 0|1 and 2|3, xor 16 joins the pairs
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 8); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 8);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 8); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 8);
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
            if (rq == 0) *reinterpret_cast<float4*>(a.agg + (size_t)bdst[it.dst_row] * 192 + it.col) = acc;
          }
        }
        // request the x rows of this warp's first item of the next pass (of the next tile after the last pass)
        if (pass < 2 || j + 1 < T) {
          const int jn = (pass < 2) ? j : j + 1, pn = (pass < 2) ? pass + 1 : 0;
          const Item nx = item_of(gwarp, jn, pn);
          load_x(nx, nx.first, jn);
        }
        group_sync(1, CF_ATHREADS);   // every warp is done with the half-tile
      }
    }
  } else if (warp >= CF_WARP_L) {
    // ================================================================== L: operand loader
    const int quad = warp & 3;
    const int my_row = quad * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const int ltid = tid - CF_WARP_L * 32;
    for (int j = 0; j < T; ++j) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int64_t r = row0 + my_row;
      const bool valid = r < cta_end;
      if (ltid == 0) {   // the block that completes the tile after next -> L2 (a tile spans two consecutive 64 KB blocks)
        for (int64_t blk = (row0 >> 7) + (j == 0 ? 1 : 3); blk <= (row0 >> 7) + 3; ++blk)
          if (blk < a.g2h_blocks && blk * TM < cta_end) prefetch_l2_bulk(a.g2h + blk * (TM * 32), TM * 32 * 16);
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {   // K quarters: hi / lo' words [16 ch, 16 ch + 16) - 8 x LDG.128 in flight per thread
        uint4 pre[8];
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            pre[q] = ldg_stream(a.g2h + g2h_index(r, 4 * ch + q));
            pre[4 + q] = ldg_stream(a.g2h + g2h_index(r, 16 + 4 * ch + q));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) pre[q] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (ch == 0 && j > 0) {   // layer 1 of the previous tile has read the operand columns
          mbar_wait(&bars[B_A1_FREE], static_cast<uint32_t>(j - 1) & 1u);
          fence_after_sync();
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          hi[4 * q + 0] = pre[q].x; hi[4 * q + 1] = pre[q].y; hi[4 * q + 2] = pre[q].z; hi[4 * q + 3] = pre[q].w;
          lo[4 * q + 0] = pre[4 + q].x; lo[4 * q + 1] = pre[4 + q].y; lo[4 * q + 2] = pre[4 + q].z; lo[4 * q + 3] = pre[4 + q].w;
        }
        tmem_st16(trow + CFC_A1HI + ch * 16, hi);
        tmem_st16(trow + CFC_A1LO + ch * 16, lo);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive(&bars[B_A1_FULL]);
    }
  } else if (T > 0) {
    // ================================================================== M: weights + MMA issue (warp-uniform, elect inside)
    if (lane == 0) {
      mbar_expect_tx(&bars[B_W], CF_W1X + CF_W1Y + CF_W2X + CF_W2Y);
      const uint8_t* p;
      p = reinterpret_cast<const uint8_t*>(a.W1x);
      for (uint32_t off = 0; off < CF_W1X; off += 16384) bulk_g2s(w1x + off, p + off, 16384, &bars[B_W]);
      p = reinterpret_cast<const uint8_t*>(a.W1y);
      for (uint32_t off = 0; off < CF_W1Y; off += 16384) bulk_g2s(w1y + off, p + off, 16384, &bars[B_W]);
      p = reinterpret_cast<const uint8_t*>(a.W2x);
      for (uint32_t off = 0; off < CF_W2X; off += 16384) bulk_g2s(w2x + off, p + off, 16384, &bars[B_W]);
      p = reinterpret_cast<const uint8_t*>(a.W2y);
      for (uint32_t off = 0; off < CF_W2Y; off += 16384) bulk_g2s(w2y + off, p + off, 16384, &bars[B_W]);
    }
    __syncwarp();
    mbar_wait(&bars[B_W], 0);
    const uint32_t s1x = smem_u32(w1x), s1y = smem_u32(w1y), s2x = smem_u32(w2x), s2y = smem_u32(w2y);
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      mbar_wait(&bars[B_A1_FULL], ph);
      fence_after_sync();
      cf_issue<HID, 128, false>(tmem + CFC_X1, tmem + CFC_A1HI, tmem + CFC_A1LO, s1x, scaled);
      mma_commit_e(&bars[B_D1X]);
      cf_issue<HID, 64, false>(tmem + CFC_Y1, tmem + CFC_A1HI, tmem + CFC_A1LO, s1y, scaled);
      mma_commit_e(&bars[B_D1Y]);
      mma_commit_e(&bars[B_A1_FREE]);
      mbar_wait(&bars[B_A2X], ph);
      if (j > 0) mbar_wait(&bars[B_X2_FREE], ph ^ 1u);
      fence_after_sync();
      cf_issue<128, 128, true>(tmem + CFC_X2, tmem + CFC_X1, 0u, s2x, scaled);
      mma_commit_e(&bars[B_D2X]);
      mbar_wait(&bars[B_A2Y], ph);
      if (j > 0) mbar_wait(&bars[B_Y2_FREE], ph ^ 1u);
      fence_after_sync();
      cf_issue<64, 64, true>(tmem + CFC_Y2, tmem + CFC_Y1, 0u, s2y, scaled);
      mma_commit_e(&bars[B_D2Y]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

void launch_cfconv_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& mw, int blk) {
  const BlkW& w = mw.blk[blk];
  CfArgs a{};
  a.W1x = reinterpret_cast<const uint32_t*>(w.hF1a); a.W2x = reinterpret_cast<const uint32_t*>(w.hF2a);
  a.W1y = reinterpret_cast<const uint32_t*>(w.hF1b); a.W2y = reinterpret_cast<const uint32_t*>(w.hF2b);
  a.b1x = w.f1ab; a.b2x = w.f2ab; a.b1y = w.f1bb; a.b2y = w.f2bb;
  a.beta_x = w.sc + 0; a.beta_y = w.sc + 1;
  a.wsc = w.hsc;
  const size_t stride = (size_t)(b.cap > 0 ? b.cap : 1);
  a.cwx = b.cw_all + (size_t)(2 * blk) * stride;
  a.cwy = b.cw_all + (size_t)(2 * blk + 1) * stride;
  a.n_rows_dev = b.counters;
  a.g2h = b.g2h;
  a.g2h_blocks = (int64_t)((stride + TM - 1) / TM);
  a.filt = b.filt;
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  a.debug_filt = c.f16_debug_filt;
  a.xcat = b.xcat;
  a.agg = b.agg;
  a.e_src = b.e_src;
  a.e_dst = b.e_dst;
  a.in_ptr = b.in_ptr;
  // (atoms without in-edges: their agg rows stay unwritten, tc_node16_kernel reads them as zero via in_ptr)
  int64_t tiles = (b.cap + TM - 1) / TM;
  const int grid = (int)(tiles < c.num_sms ? (tiles < 1 ? 1 : tiles) : c.num_sms);
  tc_cfconv_kernel<<<grid, CF_THREADS, CF_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.cfconv_f16");
}

void set_tc_cfconv_attributes() {
  cudaFuncSetAttribute(tc_cfconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF_SMEM);
}

}  // namespace agd
