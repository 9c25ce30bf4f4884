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
//   warps  0- 7  E  epilogue 1 of conv1, then of conv2: accumulator -> bias + ShiftedSoftplus (SFU) -> fp16 hi/lo' split -> written
//                   back IN PLACE over the accumulator columns (a 16-column fp32 chunk becomes 8 hi + 8 lo' words = one k-block),
//                   which is then the A operand of layer 2.  Nothing else: this is the MUFU-bound stage (2 per element).
//   warps  8-11  D  drain of the layer-2 accumulators (x 1/scale + bias, x edge weight) into a ring of two 32-column
//                   shared-memory slabs (thread = tile row): six passes per tile, four for conv1 and two for conv2.
//   warps 12-15  L  operand loader: thread = tile row, g2h row (32 x LDG.128, L1 bypass) -> tcgen05.st into the layer-1 operand
//                   columns; the tile after next is pulled into L2 with cp.async.bulk.prefetch; bookkeeping of the tile for the
//                   reducers (x-row offsets, destinations, run starts) - the loader runs a tile ahead of them.
//   warps 16-25  R  CFConv aggregation from the ring, two teams of five warps: team 0 takes the even passes (slab 0), team 1 the
//                   odd ones (slab 1).  One warp per destination run, a quarter-warp per position-in-run mod 4, lanes = float4
//                   columns - the summation order of cfconv_aggregate_kernel (four round-robin partial sums per destination,
//                   combined as (s0 + s1) + (s2 + s3)), which does not depend on where tile or CTA boundaries fall, so the
//                   result equals the unfused path bit for bit.  The x rows of a run (10 per quarter-warp = 40 rows, 95 % of the
//                   runs of a drug-like radius graph; longer runs continue on demand) are gathered into registers right after
//                   the team's previous pass: the L2 latency hides behind the other team's slab period.  All tiles of a CTA
//                   pass through these warps in order; the partial sums of a run cut by a tile boundary wait in shared
//                   memory (double-buffered by tile parity).
//   warp  26     M  weights (176 KB, cp.async.bulk, once) and every tcgen05.mma: one elected lane issues a whole layer, with
//                   compile-time TMEM addresses and descriptors that are constant offsets of one uniform base.
// A lone warp retires a dependent instruction only every ~4-6 cycles, so what bounds a role is the length of its per-tile
// instruction stream; the split above keeps every stream under ~1000 instructions per 128-edge tile.
// Registers: launched at 72 per thread (896 threads); setmaxnreg moves them per warpgroup: E 56, D 48, L 56, R + M 96.  With
// 224 KB of shared memory the L1 is ~0, so a single spilled value is an L2 round trip: every role is spill-free, and values
// that would live across all role branches (shared-memory pointers) are recomputed per role (cf_sbase).
//
// TMEM (512 columns, one tile in flight, sub-tile pipelined):
//   [  0,128) A1   layer-1 operand hi | lo' (K = 128)          free again once layer 1 of both nets has completed
//   [128,256) X1   conv1: layer-1 accumulator, then (in place) layer-2 operand
//   [256,320) Y1   conv2: the same, 64 columns
//   [320,448) X2   conv1: layer-2 accumulator                   [448,512) Y2  conv2: layer-2 accumulator
// Tensor-pipe order per tile: L1x, L1y, (wait E) L2x, (wait E) L2y - while E works on X1 the pipe runs L1y, while it works on
// Y1 the pipe runs L2x; D and R drain tile j while the pipe and E are already on tile j+1.
// What was measured and NOT adopted (profiles/r02_cfconv_experiments.md): issuing layer 2 one tile late with layer 1 of conv1
// split into two 64-column halves (E and D slow each other down: 2.70 vs 2.56 ms), the loader warps draining the odd passes
// (their load latency lands in the drain path: 3.19 ms), per-tile address arrays in the reducers (spills or serialised LDS).
#include "kernels.h"
#include "tc_filter16.cuh"

namespace agd {

using namespace tc;

constexpr int CF_THREADS = 896;                    // 28 warps x 72 registers (27 used)
constexpr int CF_WARP_D = 8, CF_WARP_L = 12, CF_WARP_R = 16, CF_WARP_M = 26;
constexpr int CF_RWARPS = 5;                       // reducer warps per team; team 0 takes the even 32-column passes, team 1 the odd ones
constexpr int CF_BKS = 3 * TM + 4;                    // ints of bookkeeping per tile parity: x-row offsets [128] | destinations [128] | run starts [128 + 1] (+ pad)
constexpr int LDS_Q = 36;                          // padded row stride (floats) of a 32-column staging buffer
constexpr uint32_t CFC_A1HI = 0, CFC_A1LO = 64, CFC_X1 = 128, CFC_Y1 = 256, CFC_X2 = 320, CFC_Y2 = 448;
constexpr uint32_t CF_W1X = 2u * 128u * 128u * 2u, CF_W1Y = 2u * 128u * 64u * 2u, CF_W2X = 2u * 128u * 128u * 2u, CF_W2Y = 2u * 64u * 64u * 2u;
constexpr int CF_STEPS = 10;                       // rows in flight per quarter-warp: 4 x 10 = 40 rows cover 95 % of the runs of a drug-like radius graph (in-degree = 32..33 radius neighbours + the local edges outside that set) in one round trip
enum { B_W = 0, B_A1_FULL, B_A1_FREE, B_D1X, B_D1Y, B_A2X, B_A2Y, B_D2X, B_D2Y, B_X2_FREE, B_Y2_FREE, B_FULL0, B_FULL1, B_EMPTY0, B_EMPTY1,
       B_BK_FULL0, B_BK_FULL1, B_BK_FREE0, B_BK_FREE1, B_COUNT };

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
  unsigned long long* timing;              // diagnostics (build with -DAGD_F16_TIMING): [role][8 phases] accumulated cycles, or nullptr
  const float* xcat;
  float* agg;
  const int *e_src, *e_dst, *in_ptr;
};

constexpr size_t CF_SMEM = 1024 + CF_W1X + CF_W1Y + CF_W2X + CF_W2Y + (128 + 64 + 128 + 64) * sizeof(float) +
                           2 * TM * LDS_Q * sizeof(float) + 2 * 4 * 192 * sizeof(float) + (2 * CF_BKS + 8) * sizeof(int) + 24 * sizeof(uint64_t) + 64;

// ---- shared-memory layout as offsets from the 1024-byte-aligned base.  Every role recomputes the base (a few instructions)
// instead of inheriting pointers from the kernel prologue: values that live across all role branches are what ptxas spills
// first, and with ~4 KB of L1 left a spilled pointer costs an L2 round trip in front of every barrier wait.
constexpr uint32_t CFO_B1X = CF_W1X + CF_W1Y + CF_W2X + CF_W2Y, CFO_B1Y = CFO_B1X + 128 * 4, CFO_B2X = CFO_B1Y + 64 * 4,
                   CFO_B2Y = CFO_B2X + 128 * 4, CFO_W = CFO_B2Y + 64 * 4, CFO_CARRY = CFO_W + 2 * TM * LDS_Q * 4,
                   CFO_BK = CFO_CARRY + 2 * 4 * 192 * 4, CFO_META = CFO_BK + 2 * CF_BKS * 4, CFO_BARS = CFO_META + 8 * 4;
__device__ __forceinline__ uint32_t cf_sbase() {
  extern __shared__ uint8_t smem_raw[];
  uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  asm volatile("" : "+r"(sb));   // opaque: not merged with (and kept alive from) another role's copy
  return sb;
}
__device__ __forceinline__ uint8_t* cf_generic(uint32_t saddr) { return static_cast<uint8_t*>(__cvta_shared_to_generic(saddr)); }
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_commit_s(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- MMA issue: one elected lane of the (converged) MMA warp runs a whole layer; the other lanes skip it (elect_one: tc16_common.cuh)

struct PhaseClock {
#ifdef AGD_F16_TIMING
  long long t_prev;
  unsigned long long acc[8];
  bool on;
  __device__ __forceinline__ void start(bool enabled) {
    on = enabled;
    for (int i = 0; i < 8; ++i) acc[i] = 0;
    t_prev = on ? clock64() : 0;
  }
  __device__ __forceinline__ void tick(int phase) {
    if (on) {
      const long long t = clock64();
      acc[phase] += static_cast<unsigned long long>(t - t_prev);
      t_prev = t;
    }
  }
  __device__ __forceinline__ void flush(unsigned long long* out, int role) {
    if (on)
      for (int i = 0; i < 8; ++i) atomicAdd(out + role * 8 + i, acc[i]);
  }
#else
  __device__ __forceinline__ void start(bool) {}
  __device__ __forceinline__ void tick(int) {}
  __device__ __forceinline__ void flush(unsigned long long*, int) {}
#endif
};

template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One layer D[128 x N] = A[128 x K] . W^T as 3 x K/16 kind::f16 MMAs in the order of issue_3xf16 (cross terms first, folded
// into the main chain by scale-input-d).  All TMEM addresses are compile-time constants (the kernel owns all 512 columns, so
// the allocation starts at column 0) and the weight descriptors are constant offsets of one uniform value: nothing has to be
// moved into uniform registers per instruction.  INPLACE: the operand was written over the accumulator in 16-column chunks
// (epilogue 1): 16 fp32 columns = the 16 k values of one k-block become 8 hi words at a + 16 kb and 8 lo' words right behind.
template <int K, int N, bool INPLACE, uint32_t D, uint32_t A_HI, uint32_t A_LO>
__device__ __forceinline__ void cf_issue(uint64_t d_hi, bool scaled, uint32_t bar0, uint32_t bar1 = 0u, uint32_t bar2 = 0u) {
  constexpr uint32_t idesc = idesc_f16(N);
  constexpr uint32_t half_bytes = static_cast<uint32_t>(K) * N * 2u;
  const uint64_t d_lo = d_hi + (half_bytes >> 4);
  if (elect_one()) {
#define CF_HI_AT(kb) (INPLACE ? A_HI + 16u * (kb) : A_HI + 8u * (kb))
#define CF_LO_AT(kb) (INPLACE ? A_HI + 16u * (kb) + 8u : A_LO + 8u * (kb))
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    mma_f16_ts(D, CF_HI_AT(kb), d_lo + boff16, idesc, kb > 0 ? 1u : 0u);
    mma_f16_ts(D, CF_LO_AT(kb), d_hi + boff16, idesc, 1u);
  }
  if (scaled) mma_f16_ts_scaled<F16_LO_SHIFT>(D, CF_HI_AT(0), d_hi, idesc);
  else mma_f16_ts(D, CF_HI_AT(0), d_hi, idesc, 1u);
#pragma unroll
  for (int kb = 1; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    mma_f16_ts(D, CF_HI_AT(kb), d_hi + boff16, idesc, 1u);
  }
  // tcgen05.commit: the barriers complete once every MMA issued so far has
  mma_commit_s(bar0);
  if (bar1) mma_commit_s(bar1);
  if (bar2) mma_commit_s(bar2);
  }
  __syncwarp();
#undef CF_HI_AT
#undef CF_LO_AT
}

__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// epilogue 1 of one 16-column accumulator chunk, in place: t = SSP(D / s1 + b1) -> 8 hi words | 8 lo' words
__device__ __forceinline__ void cf_epi1_chunk(uint32_t taddr, const float* s_b, float inv1, float lo_scale, __half2& amax) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  wait_ld();
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float2 b = *reinterpret_cast<const float2*>(s_b + 2 * q);
    const float t0 = ssp_log2(fmaf(__uint_as_float(v[2 * q]), inv1, b.x));
    const float t1 = ssp_log2(fmaf(__uint_as_float(v[2 * q + 1]), inv1, b.y));
    split2_f16(t0, t1, lo_scale, hi[q], lo[q], amax);
  }
  tmem_st8(taddr, hi);
  tmem_st8(taddr + 8, lo);
}

// One aggregation item: a destination run of the tile seen by one lane.  The quarter-warp rq (= lane / 8) owns the rows whose
// position in the run is rq mod 4 - first, first + 4, ... < e - and the lane four columns of them.
struct CfItem {
  int first, e, dst_off;   // dst_off: agg row offset of the run's destination
  bool cin, cout;          // the run continues from the previous tile / into the next one
};

// (s0 + s1) + (s2 + s3) over the four quarter-warps, or - for a run that continues in the next tile - the four partial sums
// into the carry buffer
__device__ __forceinline__ void cf_finish(const CfItem& it, float4 acc, float* carry_out, float* agg_col, int rq) {
  if (it.cout) {
    *reinterpret_cast<float4*>(carry_out) = acc;
  } else {   // xor 8 pairs quarter 0|1 and 2|3, xor 16 joins the pairs
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 8); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 8);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 8); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 8);
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (rq == 0) *reinterpret_cast<float4*>(agg_col + it.dst_off) = acc;
  }
}

// acc += x[src[row]] * W[row] for row = row, row + 4, ... < e with loads on demand: the rows of a run beyond the prefetched
// ones (in-degree > 36: caller-supplied dense graphs) and the runs beyond a warp's first two (tiles of more than 14 runs:
// molecules of fewer than ~10 atoms).  Rare, so deliberately small: four gathers in flight, inlined (a call site - even a
// not-taken one - makes ptxas spill around it on the pipelined path: measured +50 % kernel time).
__device__ __forceinline__ float4 cf_reduce_rows(const float* px, const int* bsrc, const float* sw, int row, int e, float4 acc) {
#pragma unroll 1
  for (; row < e; row += 16) {
    float4 xg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = row + 4 * u;
      xg[u] = __ldg(reinterpret_cast<const float4*>(px + bsrc[rr < e ? rr : e - 1]));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = row + 4 * u;
      const float4 w = *reinterpret_cast<const float4*>(sw + (rr < e ? rr : e - 1) * LDS_Q);
      if (rr < e) {
        acc.x = fmaf(xg[u].x, w.x, acc.x);
        acc.y = fmaf(xg[u].y, w.y, acc.y);
        acc.z = fmaf(xg[u].z, w.z, acc.z);
        acc.w = fmaf(xg[u].w, w.w, acc.w);
      }
    }
  }
  return acc;
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
  float* s_W = s_b2y + 64;                                 // [2][128][LDS_Q] ring of 32-column filter slabs awaiting aggregation
  float* s_carry = s_W + 2 * TM * LDS_Q;                   // [2 tile parities][4][192] partial sums of the run cut by a tile boundary (written in tile j, read in tile j + 1)
  int* s_bk = reinterpret_cast<int*>(s_carry + 2 * 4 * 192);   // [2 tile parities][3][128] x-row offsets (src * 192) | destinations | run starts
  int* s_meta = s_bk + 2 * CF_BKS;                         // [2][4] runs in the tile, carry in, carry out, tile row where run 0 began
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_meta + 8);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = *a.n_rows_dev;
  // this CTA's contiguous row range, snapped to run (destination) boundaries so that no run crosses CTAs
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
    mbar_init(&bars[B_A2X], 256);
    mbar_init(&bars[B_A2Y], 256);
    mbar_init(&bars[B_D2X], 1);
    mbar_init(&bars[B_D2Y], 1);
    mbar_init(&bars[B_X2_FREE], 128);
    mbar_init(&bars[B_Y2_FREE], 128);
    mbar_init(&bars[B_FULL0], 128);
    mbar_init(&bars[B_FULL1], 128);
    mbar_init(&bars[B_EMPTY0], CF_RWARPS);
    mbar_init(&bars[B_EMPTY1], CF_RWARPS);
    mbar_init(&bars[B_BK_FULL0], 1);
    mbar_init(&bars[B_BK_FULL1], 1);
    mbar_init(&bars[B_BK_FREE0], 2 * CF_RWARPS);
    mbar_init(&bars[B_BK_FREE1], 2 * CF_RWARPS);
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
  const bool scaled = a.scaled != 0;
  // The kernel addresses tensor memory by compile-time column numbers: the only CTA of the SM asked for all 512 columns, so the
  // allocation starts at column 0.  Should that ever not hold, leave through the range flag (the host repeats the call on the
  // 3xTF32 kernels) instead of computing on someone else's columns.
  if (*s_tmem != 0u) {
    if (tid == 0) atomicOr(a.range_flag, 2);
    __syncthreads();
    if (warp == 0) tmem_dealloc(*s_tmem, 512);
    return;
  }
  // Register rebalancing (setmaxnreg, one instruction per warpgroup, first thing inside each role's branch so that ptxas knows
  // which budget the branch is compiled for): reducer warps 16-25 and the MMA warp take what the loader, drain and epilogue
  // warpgroups give back.
  if (warp < CF_WARP_D) {
    reg_dec<56>();
    // ================================================================== E: epilogue 1 of conv1, then of conv2, in place
    const uint32_t sb = cf_sbase(), bar0 = sb + CFO_BARS;
    const float* s_b1x = reinterpret_cast<const float*>(cf_generic(sb + CFO_B1X));
    const float* s_b1y = s_b1x + 128;
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t trow = static_cast<uint32_t>(quad * 32) << 16;
    const float inv1x = __ldg(a.wsc + 0) * (__ldg(a.beta_x) * 1.4426950408889634f);
    const float inv1y = __ldg(a.wsc + 2) * (__ldg(a.beta_y) * 1.4426950408889634f);
    const float lo_scale = scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    PhaseClock pc;
    pc.start(a.timing != nullptr && tid == 0);
    for (int j = 0; j < T; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      mbar_wait_s(bar0 + 8 * B_D1X, ph);
      fence_after_sync();
      pc.tick(0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n0 = half * 64 + c * 16;
        if (!(a.debug_filt & 16)) cf_epi1_chunk(trow + CFC_X1 + n0, s_b1x + n0, inv1x, lo_scale, amax);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive_s(bar0 + 8 * B_A2X);
      pc.tick(1);
      mbar_wait_s(bar0 + 8 * B_D1Y, ph);
      fence_after_sync();
      pc.tick(2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n0 = half * 32 + c * 16;
        if (!(a.debug_filt & 16)) cf_epi1_chunk(trow + CFC_Y1 + n0, s_b1y + n0, inv1y, lo_scale, amax);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive_s(bar0 + 8 * B_A2Y);
      pc.tick(3);
    }
    pc.flush(a.timing, 0);
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
  } else if (warp < CF_WARP_L) {
    reg_dec<48>();
    // ================================================================== D: drain of layer 2 -> staging ring
    const uint32_t sb = cf_sbase(), bar0 = sb + CFO_BARS;
    const float* s_b2x = reinterpret_cast<const float*>(cf_generic(sb + CFO_B2X));
    const float* s_b2y = s_b2x + 128;
    float* s_W = reinterpret_cast<float*>(cf_generic(sb + CFO_W));
    const int quad = warp & 3;
    const int my_row = quad * 32 + lane;
    const uint32_t trow = static_cast<uint32_t>(quad * 32) << 16;
    const float inv2x = __ldg(a.wsc + 1), inv2y = __ldg(a.wsc + 3);
    PhaseClock pc;
    pc.start(a.timing != nullptr && tid == CF_WARP_D * 32);
    uint32_t n_pass = 0;
    for (int j = 0; j < T; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      const int64_t r = row0 + my_row;
      const bool valid = my_row < n_valid;
      float cwx = 0.f, cwy = 0.f;
      if (valid) {
        cwx = __ldg(a.cwx + r);
        cwy = __ldg(a.cwy + r);
      }
      pc.tick(0);
      // six 32-column passes: conv1 columns [0,128) (passes 0-3), conv2 columns [0,64) = xcat / agg columns 128..191 (passes 4, 5)
#pragma unroll 1
      for (int pass = 0; pass < 6; ++pass, ++n_pass) {
        const bool isy = pass >= 4;
        const int n0 = isy ? (pass - 4) * 32 : pass * 32;
        const uint32_t b = n_pass & 1u, use = n_pass >> 1;
        if (use > 0) mbar_wait_s(bar0 + 8 * (B_EMPTY0 + b), (use - 1) & 1u);   // the reducers are done with this slab's previous content
        pc.tick(1);
        if (pass == 0) mbar_wait_s(bar0 + 8 * B_D2X, ph);
        if (pass == 4) mbar_wait_s(bar0 + 8 * B_D2Y, ph);
        fence_after_sync();
        pc.tick(2);
        const float inv2 = isy ? inv2y : inv2x, cw = isy ? cwy : cwx;
        const float4* sb4 = reinterpret_cast<const float4*>((isy ? s_b2y : s_b2x) + n0);
        float4* dstW = reinterpret_cast<float4*>(s_W + b * (TM * LDS_Q) + my_row * LDS_Q);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[16];
          if (a.debug_filt & 8) {
            if (cc == 1 && (pass == 3 || pass == 5)) mbar_arrive_s(bar0 + 8 * (isy ? B_Y2_FREE : B_X2_FREE));
            continue;
          }
          tmem_ld16(trow + (isy ? CFC_Y2 : CFC_X2) + n0 + 16 * cc, v);
          wait_ld();
          if (cc == 1 && (pass == 3 || pass == 5)) {   // the accumulator of this net is drained: layer 2 of the next tile may overwrite it
            fence_before_sync();
            mbar_arrive_s(bar0 + 8 * (isy ? B_Y2_FREE : B_X2_FREE));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = sb4[4 * cc + q];
            float4 o;
            o.x = fmaf(__uint_as_float(v[q * 4 + 0]), inv2, bq.x) * cw;
            o.y = fmaf(__uint_as_float(v[q * 4 + 1]), inv2, bq.y) * cw;
            o.z = fmaf(__uint_as_float(v[q * 4 + 2]), inv2, bq.z) * cw;
            o.w = fmaf(__uint_as_float(v[q * 4 + 3]), inv2, bq.w) * cw;
            dstW[4 * cc + q] = o;
          }
        }
        if ((a.debug_filt & 1) && valid) {   // tests: the filter tensor as the unfused path writes it
#pragma unroll
          for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(a.filt + r * 192 + 32 * pass + 4 * q) = dstW[q];
        }
        mbar_arrive_s(bar0 + 8 * (B_FULL0 + b));
        pc.tick(3);
      }
    }
    pc.flush(a.timing, 1);
  } else if (warp >= CF_WARP_R) {
    reg_inc<96>();
    if (warp < CF_WARP_M) {
    // ================================================================== R: CFConv aggregation from the staging ring
    // Runs rw, rw + 7, rw + 14, ... of the tile belong to reducer warp rw; quarter-warp rq takes the rows whose position in the
    // run is rq mod 4, in order; lanes = float4 columns of the 32-column pass.  Summation order == cfconv_aggregate_kernel
    // (schnet.cu).  The x rows of the warp's first two runs are requested one pass ahead; steps past the end of a run carry
    // x = 0 and the W of the run's last row, so the unconditional fmaf leaves the sum unchanged.
    const uint32_t sb = cf_sbase(), bar0 = sb + CFO_BARS;
    float* s_W = reinterpret_cast<float*>(cf_generic(sb + CFO_W));
    float* s_carry = s_W + 2 * TM * LDS_Q;
    int* s_bk = reinterpret_cast<int*>(s_carry + 2 * 4 * 192);
    int* s_meta = s_bk + 2 * CF_BKS;
    const int team = (warp - CF_WARP_R) / CF_RWARPS, rw = (warp - CF_WARP_R) % CF_RWARPS;
    const int rq = lane >> 3, c4 = (lane & 7) * 4;
    auto item_of = [&](int k, int jt) {
      const int* bdst_ = s_bk + (jt & 1) * CF_BKS + TM;
      const int* bruns = bdst_ + TM;   // run starts, closed by the number of valid rows
      const int* meta = s_meta + (jt & 1) * 4;
      const int n_runs = meta[0];
      CfItem it;
      if (k < n_runs) {
        const int s = bruns[k];
        it.e = bruns[k + 1];
        const int base = (k == 0) ? meta[3] : s;
        it.first = s + ((rq - (s - base)) & 3);
        it.dst_off = bdst_[s] * 192;
        it.cin = (k == 0) && meta[1] != 0;
        it.cout = (k == n_runs - 1) && meta[2] != 0;
      } else {
        it.first = it.e = it.dst_off = 0;
        it.cin = it.cout = false;
      }
      return it;
    };
    // x values of the rows first + 4 u of this warp's run for its next pass.  A team only works on every other pass, so a request
    // issued right after a pass has a whole slab period of the other team to arrive before it is consumed.
    float4 xv[CF_STEPS];
    CfItem cur;   // this warp's run of the current tile (replaced by the next tile's once the team's last pass is done)
    cur.first = cur.e = cur.dst_off = 0;
    cur.cin = cur.cout = false;
    PhaseClock pc;
    pc.start(a.timing != nullptr && lane == 0 && rw == 0 && team == 0);
    const bool skip_gather = (a.debug_filt & 2) != 0;
    // (a macro, not a lambda: called from three places, a lambda is not inlined and xv would live in local memory)
#define CF_REQUEST(it, bsrc_, col_)                                                                                  \
  {                                                                                                                  \
    const float* px_ = a.xcat + (col_);                                                                              \
    int first_ = (it).first, e_ = skip_gather ? 0 : (it).e;                                                          \
    asm volatile("" : "+r"(first_), "+r"(e_)); /* opaque: keeps ptxas from parking pass-invariant indices on the stack */ \
    _Pragma("unroll") for (int u = 0; u < CF_STEPS; ++u) {                                                           \
      const int rr = first_ + 4 * u;                                                                                 \
      if (rr < e_) xv[u] = __ldg(reinterpret_cast<const float4*>(px_ + (bsrc_)[rr]));                                \
    }                                                                                                                \
  }
  // Steps past the end of the run are never loaded, so their x stays what CF_ZERO_X made it when the warp took the run over:
  // zero once per tile instead of once per pass (and never a stale value of another run, i.e. of another molecule).
#define CF_ZERO_X() \
  { _Pragma("unroll") for (int u = 0; u < CF_STEPS; ++u) xv[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const uint32_t full = bar0 + 8 * (B_FULL0 + team), empty = bar0 + 8 * (B_EMPTY0 + team);
    const float* sw = s_W + team * (TM * LDS_Q) + c4;   // the team's slab
    uint32_t n_use = 0;
    if (T > 0) {
      mbar_wait_s(bar0 + 8 * B_BK_FULL0, 0u);
      cur = item_of(rw, 0);
      CF_ZERO_X();
      CF_REQUEST(cur, s_bk, 32 * team + c4);
    }
    for (int j = 0; j < T; ++j) {
      const int n_runs = s_meta[(j & 1) * 4];
      const int* bsrc = s_bk + (j & 1) * CF_BKS;
      const bool more = j + 1 < T;
      pc.tick(0);
#pragma unroll 1
      for (int pass = team; pass < 6; pass += 2, ++n_use) {
        const int col = 32 * pass + c4;
        float* carry_w = s_carry + (j & 1) * 768 + rq * 192 + col;              // written in this tile, read in the next
        const float* carry_r = s_carry + ((j & 1) ^ 1) * 768 + rq * 192 + col;
        mbar_wait_s(full, n_use & 1u);
        pc.tick(1);
        // (opaque copy: evaluated before the wait, the condition below becomes a predicate that ptxas saves to the stack across
        // the wait loop and reloads right behind it - an L2 round trip in front of every consume: 300 k local loads per launch)
        int n_runs_ = n_runs;
        asm volatile("" : "+r"(n_runs_));
        if (rw < n_runs_ && !(a.debug_filt & 4)) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cur.cin) acc = *reinterpret_cast<const float4*>(carry_r);
          // (opaque copies: the twelve row indices below are pass-invariant, and ptxas would rather keep them on the stack across
          // the pass loop - i.e. reload them through L2 every pass - than recompute them)
          int first_ = cur.first, last_ = cur.e - 1;
          asm volatile("" : "+r"(first_), "+r"(last_));
#pragma unroll
          for (int u = 0; u < CF_STEPS; ++u) {   // branch-free: past the run x is 0 (request) and W is re-read from the run's last row
            const int rr = first_ + 4 * u;
            const float4 w = *reinterpret_cast<const float4*>(sw + (rr <= last_ ? rr : last_) * LDS_Q);
            acc.x = fmaf(xv[u].x, w.x, acc.x);
            acc.y = fmaf(xv[u].y, w.y, acc.y);
            acc.z = fmaf(xv[u].z, w.z, acc.z);
            acc.w = fmaf(xv[u].w, w.w, acc.w);
          }
          if (cur.first + 4 * CF_STEPS < cur.e) acc = cf_reduce_rows(a.xcat + col, bsrc, sw, cur.first + 4 * CF_STEPS, cur.e, acc);
          cf_finish(cur, acc, carry_w, a.agg + col, rq);
          for (int k = rw + CF_RWARPS; k < n_runs; k += CF_RWARPS) {   // more than 7 runs in the tile (short runs): on demand
            const CfItem it = item_of(k, j);
            const float4 a2 = cf_reduce_rows(a.xcat + col, bsrc, sw, it.first, it.e, make_float4(0.f, 0.f, 0.f, 0.f));
            cf_finish(it, a2, carry_w, a.agg + col, rq);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(empty);   // the slab may be refilled
        pc.tick(2);
        if (pass + 2 < 6) {
          CF_REQUEST(cur, bsrc, col + 64);
        } else if (more) {   // the next tile's run: its bookkeeping is published by the loader, which runs a tile ahead
          mbar_wait_s(bar0 + 8 * (B_BK_FULL0 + ((j + 1) & 1)), static_cast<uint32_t>((j + 1) >> 1) & 1u);
          cur = item_of(rw, j + 1);
          const int* bsrc_n = s_bk + ((j + 1) & 1) * CF_BKS;
          CF_ZERO_X();
          CF_REQUEST(cur, bsrc_n, 32 * team + c4);
        }
        pc.tick(3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_s(bar0 + 8 * (B_BK_FREE0 + (j & 1)));
    }
#undef CF_REQUEST
#undef CF_ZERO_X
    pc.flush(a.timing, 2);
    } else if (warp == CF_WARP_M && T > 0) {
    // ================================================================== M: weights + MMA issue (warp-uniform, elect inside)
    const uint32_t sb = cf_sbase(), bar0 = sb + CFO_BARS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(cf_generic(bar0));
    uint8_t* w1x = cf_generic(sb);
    uint8_t* w1y = w1x + CF_W1X;
    uint8_t* w2x = w1y + CF_W1Y;
    uint8_t* w2y = w2x + CF_W2X;
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
    mbar_wait_s(bar0 + 8 * B_W, 0);
    // weight descriptors: the four [hi | lo'] images follow each other from `base`, so each is the first plus a constant
    const uint64_t d1x = smem_desc_sw128(sb);
    const uint64_t d1y = d1x + (CF_W1X >> 4), d2x = d1y + (CF_W1Y >> 4), d2y = d2x + (CF_W2X >> 4);
    PhaseClock pc;
    pc.start(a.timing != nullptr && lane == 0);
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
      const uint32_t ph = static_cast<uint32_t>(j) & 1u;
      mbar_wait_s(bar0 + 8 * B_A1_FULL, ph);
      fence_after_sync();
      pc.tick(0);
      cf_issue<HID, 128, false, CFC_X1, CFC_A1HI, CFC_A1LO>(d1x, scaled, bar0 + 8 * B_D1X);
      cf_issue<HID, 64, false, CFC_Y1, CFC_A1HI, CFC_A1LO>(d1y, scaled, bar0 + 8 * B_D1Y, bar0 + 8 * B_A1_FREE);
      pc.tick(1);
      mbar_wait_s(bar0 + 8 * B_A2X, ph);
      pc.tick(2);
      if (j > 0) mbar_wait_s(bar0 + 8 * B_X2_FREE, ph ^ 1u);
      fence_after_sync();
      pc.tick(3);
      cf_issue<128, 128, true, CFC_X2, CFC_X1, 0u>(d2x, scaled, bar0 + 8 * B_D2X);
      pc.tick(4);
      mbar_wait_s(bar0 + 8 * B_A2Y, ph);
      pc.tick(5);
      if (j > 0) mbar_wait_s(bar0 + 8 * B_Y2_FREE, ph ^ 1u);
      fence_after_sync();
      pc.tick(6);
      cf_issue<64, 64, true, CFC_Y2, CFC_Y1, 0u>(d2y, scaled, bar0 + 8 * B_D2Y);
      pc.tick(7);
    }
    pc.flush(a.timing, 4);
    }
  } else {
    reg_dec<56>();
    // ================================================================== L: operand loader + bookkeeping
    const uint32_t sb = cf_sbase(), bar0 = sb + CFO_BARS;
    int* s_bk = reinterpret_cast<int*>(cf_generic(sb + CFO_BK));
    int* s_meta = s_bk + 2 * CF_BKS;
    const int quad = warp & 3;
    const int my_row = quad * 32 + lane;
    const uint32_t trow = static_cast<uint32_t>(quad * 32) << 16;
    const int ltid = tid - CF_WARP_L * 32;
    PhaseClock pc;
    pc.start(a.timing != nullptr && ltid == 0);
    // Bookkeeping of tile jn -> s_bk[jn & 1]: x-row offsets and destinations of the rows, run starts, and whether the first / last
    // run continues across the tile boundary.  Published through B_BK_FULL; the buffer was released by the reducers two tiles ago.
    auto bookkeep = [&](int jn) {
      int* bsrc = s_bk + (jn & 1) * CF_BKS;
      int* bdst = bsrc + TM;
      int* bruns = bdst + TM;
      const int64_t rown = cta_begin + static_cast<int64_t>(jn) * TM;
      const int nv = (cta_end - rown < TM) ? static_cast<int>(cta_end - rown) : TM;
      const int src = (my_row < nv) ? __ldg(a.e_src + rown + my_row) * 192 : 0;
      const int dst = (my_row < nv) ? __ldg(a.e_dst + rown + my_row) : -1;
      int prev_dst = -2, next_dst = -5;
      if (quad == 0) {
        if (jn > 0) prev_dst = __ldg(a.e_dst + rown - 1);
        if (rown + nv < cta_end) next_dst = __ldg(a.e_dst + rown + nv);
      }
      if (jn >= 2) mbar_wait_s(bar0 + 8 * (B_BK_FREE0 + (jn & 1)), static_cast<uint32_t>((jn >> 1) - 1) & 1u);
      bsrc[my_row] = src;
      bdst[my_row] = dst;
      group_sync(2, 128);
      if (quad == 0) {
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
        if (lane == 0) bruns[n_runs] = nv;   // closes the last run
        const int d0 = bdst[0];
        const bool cin = (d0 == prev_dst);
        // tile row at which run 0 began (<= 0: in an earlier tile): positions in a run are counted from its first edge
        const int base0 = cin ? static_cast<int>(static_cast<int64_t>(__ldg(a.in_ptr + d0)) - rown) : 0;
        if (lane == 0) {
          int* meta = s_meta + (jn & 1) * 4;
          meta[0] = n_runs;
          meta[1] = cin ? 1 : 0;
          meta[2] = (next_dst == bdst[nv - 1]) ? 1 : 0;
          meta[3] = base0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(bar0 + 8 * (B_BK_FULL0 + (jn & 1)));
      }
    };
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
        if (valid && !(a.debug_filt & 32)) {
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
          pc.tick(0);
          mbar_wait_s(bar0 + 8 * B_A1_FREE, static_cast<uint32_t>(j - 1) & 1u);
          fence_after_sync();
          pc.tick(1);
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
      mbar_arrive_s(bar0 + 8 * B_A1_FULL);
      pc.tick(2);
      bookkeep(j);   // the loader runs a tile ahead of the reducers: never urgent
      pc.tick(3);
    }
    pc.flush(a.timing, 3);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(0u, 512);
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
  a.timing = c.f16_timing;
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
