// CFConv filter network on tcgen05 with fp16-split operands, TWO edge tiles in flight per SM (schnet.py:136-162, merged
// as in pack.py):
//
//   W_e = ( F2 . SSP_beta( F1 . g2_e + b1 ) + b2 ) * cw_e
//
// Numerics ("3xFP16"): fp16 carries the same 11 significant bits as TF32, so x = hi + lo with hi = rn_f16(x) and
// lo' = rn_f16((x - hi) * 2^S) reproduces the 3xTF32 split (2^-22 relative) at TWICE the tensor-core rate and HALF the
// operand footprint.  The cross terms are accumulated first at scale 2^S and folded into the main chain by the
// instruction's scale-input-d (D = A.B + D * 2^-S), so neither part ever leaves the fp16 normal range:
//
//   D  = sum_k ( a_hi . w_lo' + a_lo' . w_hi )            (scale 2^S)
//   D  = a_hi[0] . w_hi[0] + D * 2^-S ;  D += a_hi[k] . w_hi[k]  (k >= 1)
//
// Weights are pre-scaled per matrix by a power of two (pack.umma_image_f16) so that max|w| sits at 2^13..2^14; the
// epilogue multiplies by the inverse (exact).  Activations beyond the fp16 range (|x| > 65000) raise a device flag and
// the host re-runs the call on the 3xTF32 kernels (tc_filter.cu) - never silently wrong.
//
// Pipeline (one CTA per SM, 512 threads, 512 TMEM columns = 2 slots x [D 128 | A_hi 64 | A_lo 64]):
//   * both layers' weight images (hi | lo', K-major SWIZZLE_128B, <= 128 KB) are bulk-copied into shared memory ONCE;
//   * warps 0-7 own slot 0, warps 8-15 slot 1: each group stages its tile's operand rows into TMEM (tcgen05.st), runs
//     the SSP epilogue TMEM -> registers -> TMEM and the output epilogue TMEM -> HBM, and signals "operand ready" on an
//     mbarrier; thread (quad q, lane l, half h) owns tile row 32q+l and one half of the columns;
//   * the group's first thread is its MMA issuer: once all 256 threads have arrived it issues the layer's 3 x K/16
//     tcgen05.mma.kind::f16 (M=128, N=F, K=16, A from TMEM) and commits to the slot's "accumulator ready" mbarrier.
//     While one group runs an epilogue the tensor core works on the other slot.
#include <cuda_fp16.h>

#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace agd {

namespace tc {
// kind::f16 instruction descriptor: fp32 accumulate, F16 x F16, both K-major, M=128
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// D = A.B + D * 2^-SHIFT
template <int SHIFT>
__device__ __forceinline__ void mma_f16_ts_scaled(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p, %9;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(1u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "n"(SHIFT)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr)
      : "memory");
}
}  // namespace tc

constexpr int F16_THREADS = 512;        // 2 groups x 8 warps
constexpr int F16_GROUP = 256;
constexpr int F16_GWARPS = 8;
constexpr int SLOT_COLS = 256;          // TMEM columns per slot
constexpr int C16_D = 0, C16_AHI = 128, C16_ALO = 192;
constexpr int LDS_W = 68;               // padded row stride (floats) of the 64-column filter half-tile awaiting aggregation

struct TcF16Args {
  const uint32_t* W1img;   // [hi | lo'] fp16 images of F1 (K=128): each (128/64) x F rows x 128 B
  const uint32_t* W2img;   // ... of F2 (K=F)
  const float *f1b, *f2b, *dw, *beta_ptr;
  const float* wsc;        // [0] = 1/scale(F1), [1] = 1/scale(F2)
  const int* n_rows_dev;
  const uint4* g2h;        // pre-split encoder state (tc_common.cuh: g2h_index)
  const float* e_len;
  float* filt;             // [E][192]  (!FUSE)
  int col0;                // 0 (conv1) or 128 (conv2): column offset in filt / xcat / agg
  float cutoff;
  int smooth;
  int scaled;              // 1: lo' scaled by 2^S + scale-input-d, 0: unscaled lo (A/B switch AGD_F16_LOSHIFT=0)
  int* range_flag;
  int debug_filt;          // FUSE: also write the filter tensor (tests / diagnostics)
  // fused aggregation (FUSE): agg[dst][col0 + n] = sum over the destination's edges, in CSC order, of x[src][col0 + n] * W_e[n]
  const float* xcat;       // [N][192]
  float* agg;              // [N][192], zeroed before the launch
  const int *e_src, *e_dst, *in_ptr;
};

template <int F, bool FUSE>
struct TcF16Smem {
  static constexpr uint32_t W1_HALF = 128u * F * 2u, W2_HALF = static_cast<uint32_t>(F) * F * 2u;
  static constexpr size_t fuse_bytes = FUSE ? (2 * TM * LDS_W + 2 * 128) * sizeof(float) + 6 * TM * sizeof(int) : 0;
  static constexpr size_t bytes = 1024 + 2 * W1_HALF + 2 * W2_HALF + (128 + 128 + 128 + 256) * sizeof(float) + fuse_bytes +
                                  16 * sizeof(uint64_t) + 64;
};

// the 3 x K/16 MMAs of one layer for one slot, issued by ONE thread: cross terms first, then the main chain
template <int K, int N>
__device__ __forceinline__ void issue_3xf16(uint32_t slot, uint32_t w_smem, uint32_t half_bytes, bool scaled) {
  constexpr uint32_t idesc = tc::idesc_f16(N);
  const uint64_t d_hi = tc::smem_desc_sw128(w_smem), d_lo = tc::smem_desc_sw128(w_smem + half_bytes);
  const uint32_t D = slot + C16_D, ahi = slot + C16_AHI, alo = slot + C16_ALO;
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    tc::mma_f16_ts(D, ahi + kb * 8, d_lo + boff16, idesc, kb > 0 ? 1u : 0u);
    tc::mma_f16_ts(D, alo + kb * 8, d_hi + boff16, idesc, 1u);
  }
  if (scaled) tc::mma_f16_ts_scaled<F16_LO_SHIFT>(D, ahi, d_hi, idesc);
  else tc::mma_f16_ts(D, ahi, d_hi, idesc, 1u);
#pragma unroll
  for (int kb = 1; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    tc::mma_f16_ts(D, ahi + kb * 8, d_hi + boff16, idesc, 1u);
  }
}

// Fused CFConv aggregation of one 64-column half-tile staged in shared memory (FUSE).  Edges are CSC-sorted, so the rows of one
// destination ("run") are consecutive; warp w of the group reduces runs w, w+8, ... with lanes = column pairs, walking
// the run's rows IN ORDER with one fmaf per row - the same arithmetic, in the same order, as cfconv_aggregate_kernel, so
// the sum for a destination does not depend on where tile or CTA boundaries fall.  A run cut by a tile boundary is
// continued, not re-associated: the partial sum travels to the other slot's group (which owns the next tile) through
// `carry` guarded by a full/empty mbarrier pair.  CTA ranges start at run boundaries, so nothing crosses CTAs.
struct RunCtx {
  int n_runs, n_valid;
  bool carry_in;        // run 0 continues the previous tile's last run
  bool carry_out;       // the last run continues in the next tile
};

template <int F>
__device__ __forceinline__ void aggregate_half(const TcF16Args& a, const RunCtx& rc, const float* s_W, const int* s_src,
                                               const int* s_dst, const int* s_runs, int pass, float* carry_mine,
                                               const float* carry_other, uint64_t* full_mine, uint64_t* empty_mine,
                                               uint64_t* full_other, uint64_t* empty_other, uint32_t prod_cnt, uint32_t cons_cnt,
                                               int gwarp, int lane) {
  constexpr int PASSES = F / 64;
  const int colg = a.col0 + pass * 64 + 2 * lane;   // this lane's two columns in xcat / agg
#pragma unroll 1
  for (int k = gwarp; k < rc.n_runs; k += F16_GWARPS) {
    const int s = s_runs[k];
    const int e = (k + 1 < rc.n_runs) ? s_runs[k + 1] : rc.n_valid;
    float2 acc = make_float2(0.f, 0.f);
    if (k == 0 && rc.carry_in) {
      if (pass == 0) tc::mbar_wait(full_other, cons_cnt & 1u);
      acc = *reinterpret_cast<const float2*>(carry_other + pass * 64 + 2 * lane);
      if (pass == PASSES - 1) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(empty_other);
      }
    }
    int row = s;
    for (; row + 4 <= e; row += 4) {
      float2 xv[4], wv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xv[u] = __ldg(reinterpret_cast<const float2*>(a.xcat + (size_t)s_src[row + u] * 192 + colg));
        wv[u] = *reinterpret_cast<const float2*>(s_W + (row + u) * LDS_W + 2 * lane);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(xv[u].x, wv[u].x, acc.x);
        acc.y = fmaf(xv[u].y, wv[u].y, acc.y);
      }
    }
    for (; row < e; ++row) {
      const float2 xv = __ldg(reinterpret_cast<const float2*>(a.xcat + (size_t)s_src[row] * 192 + colg));
      const float2 wv = *reinterpret_cast<const float2*>(s_W + row * LDS_W + 2 * lane);
      acc.x = fmaf(xv.x, wv.x, acc.x);
      acc.y = fmaf(xv.y, wv.y, acc.y);
    }
    if (k == rc.n_runs - 1 && rc.carry_out) {
      if (pass == 0) tc::mbar_wait(empty_mine, (prod_cnt & 1u) ^ 1u);   // the previous carry of this group was consumed
      *reinterpret_cast<float2*>(carry_mine + pass * 64 + 2 * lane) = acc;
      if (pass == PASSES - 1) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(full_mine);
      }
    } else {
      *reinterpret_cast<float2*>(a.agg + (size_t)s_dst[s] * 192 + colg) = acc;
    }
  }
}

template <int F, bool FUSE>
__global__ void __launch_bounds__(F16_THREADS, 1) tc_filter16_kernel(const TcF16Args a) {
  using namespace tc;
  using SM = TcF16Smem<F, FUSE>;
  constexpr uint32_t W1_HALF = SM::W1_HALF, W2_HALF = SM::W2_HALF;
  constexpr int HC = F / 2;   // output columns owned by one thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w1 = base;
  uint8_t* w2 = base + 2 * W1_HALF;
  float* s_b1 = reinterpret_cast<float*>(w2 + 2 * W2_HALF);
  float* s_b2 = s_b1 + 128;
  float* s_dw = s_b2 + 128;
  float* s_cw = s_dw + 128;                                        // [2][128] envelope weight of each slot's rows
  float* s_Wt = s_cw + 256;                                        // FUSE: [2][128][LDS_W] filter half-tiles
  float* s_carry = s_Wt + (FUSE ? 2 * TM * LDS_W : 0);             // FUSE: [2][128] partial sums of runs cut by a tile boundary
  int* s_srcdst = reinterpret_cast<int*>(s_carry + (FUSE ? 256 : 0));   // FUSE: [2][3][128] src / dst / run starts of each slot's rows
  // [0] weights landed, [1+g] operand ready, [3+g] accumulator ready, [5+g] carry full, [7+g] carry empty
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_srcdst + (FUSE ? 6 * TM : 0));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = *a.n_rows_dev;

  // this CTA's contiguous row range; with FUSE it is snapped to run (destination) boundaries
  const int64_t rows_per_cta = ((static_cast<int64_t>(n_rows) + TM - 1) / TM + gridDim.x - 1) / gridDim.x * TM;
  int64_t cta_begin = static_cast<int64_t>(blockIdx.x) * rows_per_cta, cta_end = cta_begin + rows_per_cta;
  if (cta_begin > n_rows) cta_begin = n_rows;
  if (cta_end > n_rows) cta_end = n_rows;
  if (FUSE) {
    if (cta_begin > 0 && cta_begin < n_rows) cta_begin = __ldg(a.in_ptr + __ldg(a.e_dst + cta_begin));
    if (cta_end < n_rows) cta_end = __ldg(a.in_ptr + __ldg(a.e_dst + cta_end));
  }
  const int cta_tiles = static_cast<int>((cta_end - cta_begin + TM - 1) / TM);   // <= 0: nothing to do

  if (warp == 0) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], F16_GROUP);
    mbar_init(&bars[2], F16_GROUP);
    for (int i = 3; i < 9; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (tid < F) {
    s_b1[tid] = __ldg(a.f1b + tid);
    s_b2[tid] = __ldg(a.f2b + tid);
  }
  if (tid >= 128 && tid < 256) s_dw[tid - 128] = __ldg(a.dw + tid - 128);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (tid == 0 && cta_tiles > 0) {   // both layers' weight images, once
    mbar_expect_tx(&bars[0], 2 * W1_HALF + 2 * W2_HALF);
    const uint8_t* s1 = reinterpret_cast<const uint8_t*>(a.W1img);
    const uint8_t* s2 = reinterpret_cast<const uint8_t*>(a.W2img);
    for (uint32_t off = 0; off < 2 * W1_HALF; off += 16384) bulk_g2s(w1 + off, s1 + off, 16384, &bars[0]);
    for (uint32_t off = 0; off < 2 * W2_HALF; off += 16384) bulk_g2s(w2 + off, s2 + off, 16384, &bars[0]);
  }
  {
    // ------------------------------------------------------------------ operand staging + MMA issue + epilogues of one slot
    const int g = warp >> 3, gwarp = warp & 7, quad = warp & 3, half = (warp >> 2) & 1;
    const int my_row = quad * 32 + lane;
    const uint32_t slot = tmem + static_cast<uint32_t>(g * SLOT_COLS);
    const uint32_t trow = slot + (static_cast<uint32_t>(quad * 32) << 16);
    uint64_t* a_ready = &bars[1 + g];
    uint64_t* d_ready = &bars[3 + g];
    float* cwbuf = s_cw + g * 128;
    float* s_W = s_Wt + g * TM * LDS_W;
    int* s_src = s_srcdst + g * 3 * TM;
    int* s_dst = s_src + TM;
    int* s_runs = s_dst + TM;
    const float beta = __ldg(a.beta_ptr);
    const float inv1 = __ldg(a.wsc + 0), inv2 = __ldg(a.wsc + 1);
    const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    float amax = 0.f;
    uint32_t dph = 0, aph = 0, prod_cnt = 0, cons_cnt = 0;
    const bool issuer = (tid & (F16_GROUP - 1)) == 0;
    const bool scaled = a.scaled != 0;
    if (issuer && g < cta_tiles) mbar_wait(&bars[0], 0);

    // this thread's slice of the pre-split g2 row: 32 hi words + 32 lo' words (K = 64*half .. 64*half+63), one tile ahead
    uint4 pre[16];
    float pre_len = -1.f;
    auto prefetch = [&](int j) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int64_t r = row0 + my_row;
      if (j < cta_tiles && r < cta_end) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          pre[q] = __ldg(a.g2h + g2h_index(r, 8 * half + q));
          pre[8 + q] = __ldg(a.g2h + g2h_index(r, 16 + 8 * half + q));
        }
        pre_len = __ldg(a.e_len + r);
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) pre[q] = make_uint4(0u, 0u, 0u, 0u);
        pre_len = -1.f;
      }
    };

    prefetch(g);
    for (int j = g; j < cta_tiles; j += 2) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      const int64_t r = row0 + my_row;
      const bool valid = my_row < n_valid;
      // ---- stage A: the pre-split g2 rows go straight into the slot's operand columns
      if (half == 1) cwbuf[my_row] = valid ? cfconv_edge_weight_smem(pre_len, s_dw, a.cutoff, a.smooth) : 0.f;
      if (FUSE && half == 0) {
        s_src[my_row] = valid ? __ldg(a.e_src + r) : 0;
        s_dst[my_row] = valid ? __ldg(a.e_dst + r) : -1;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          hi[4 * q + 0] = pre[4 * c + q].x; hi[4 * q + 1] = pre[4 * c + q].y; hi[4 * q + 2] = pre[4 * c + q].z; hi[4 * q + 3] = pre[4 * c + q].w;
          lo[4 * q + 0] = pre[8 + 4 * c + q].x; lo[4 * q + 1] = pre[8 + 4 * c + q].y; lo[4 * q + 2] = pre[8 + 4 * c + q].z; lo[4 * q + 3] = pre[8 + 4 * c + q].w;
        }
        tmem_st16(trow + C16_AHI + half * 32 + c * 16, hi);
        tmem_st16(trow + C16_ALO + half * 32 + c * 16, lo);
      }
      wait_st();
      fence_before_sync();
      group_sync(1 + g, F16_GROUP);
      const float cw = cwbuf[my_row];
      mbar_arrive(a_ready);
      if (issuer) {
        mbar_wait(a_ready, aph);
        aph ^= 1u;
        fence_after_sync();
        issue_3xf16<HID, F>(slot, smem_u32(w1), W1_HALF, scaled);
        mma_commit(d_ready);
      }
      // run structure of this tile (FUSE; every warp computes the same masks while layer 1 runs)
      RunCtx rc;
      if (FUSE) {
        const int prev_dst = (j > 0) ? __ldg(a.e_dst + row0 - 1) : -2;
        rc.n_valid = n_valid;
        int below = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int row = w * 32 + lane;
          const int d = s_dst[row];
          const int dp = (row > 0) ? s_dst[row - 1] : -3;
          const bool start = row < n_valid && (row == 0 || d != dp);
          const uint32_t m = __ballot_sync(0xffffffffu, start);
          if (gwarp == 0 && start) s_runs[below + __popc(m & ((1u << lane) - 1u))] = row;   // published by the next group_sync
          below += __popc(m);
        }
        rc.n_runs = below;
        rc.carry_in = (s_dst[0] == prev_dst);
        rc.carry_out = (row0 + n_valid < cta_end) && (__ldg(a.e_dst + row0 + n_valid) == s_dst[n_valid - 1]);
      }
      // ---- layer 1 done -> epilogue 1: t = SSP_beta(D / s1 + b1) -> operand (hi / lo') of layer 2
      mbar_wait(d_ready, dph);
      dph ^= 1u;
      fence_after_sync();
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) {
        const int n0 = half * HC + c * 32;
        uint32_t v[32];
        tmem_ld32(trow + C16_D + n0, v);
        wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float t0 = ssp_fast(fmaf(__uint_as_float(v[2 * q]), inv1, s_b1[n0 + 2 * q]), beta);
          const float t1 = ssp_fast(fmaf(__uint_as_float(v[2 * q + 1]), inv1, s_b1[n0 + 2 * q + 1]), beta);
          split2_f16(t0, t1, lo_scale, hi[q], lo[q], amax);
        }
        tmem_st16(trow + C16_AHI + (n0 >> 1), hi);
        tmem_st16(trow + C16_ALO + (n0 >> 1), lo);
      }
      wait_st();
      fence_before_sync();
      mbar_arrive(a_ready);
      if (issuer) {
        mbar_wait(a_ready, aph);
        aph ^= 1u;
        fence_after_sync();
        issue_3xf16<F, F>(slot, smem_u32(w2), W2_HALF, scaled);
        mma_commit(d_ready);
      }
      prefetch(j + 2);   // next tile's rows travel while layer 2 runs
      // ---- layer 2 done -> epilogue 2: W = (D / s2 + b2) * cw
      mbar_wait(d_ready, dph);
      dph ^= 1u;
      fence_after_sync();
      if (FUSE) {
        // two column halves per thread are staged as 64-column half-tiles: pass p holds filter columns [64p, 64p+64)
#pragma unroll 1
        for (int pass = 0; pass < F / 64; ++pass) {
          const int n0 = pass * 64 + half * 32;
          uint32_t v[32];
          tmem_ld32(trow + C16_D + n0, v);
          wait_ld();
          float4* dstW = reinterpret_cast<float4*>(s_W + my_row * LDS_W + half * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o;
            o.x = fmaf(__uint_as_float(v[q * 4 + 0]), inv2, s_b2[n0 + q * 4 + 0]) * cw;
            o.y = fmaf(__uint_as_float(v[q * 4 + 1]), inv2, s_b2[n0 + q * 4 + 1]) * cw;
            o.z = fmaf(__uint_as_float(v[q * 4 + 2]), inv2, s_b2[n0 + q * 4 + 2]) * cw;
            o.w = fmaf(__uint_as_float(v[q * 4 + 3]), inv2, s_b2[n0 + q * 4 + 3]) * cw;
            dstW[q] = o;
            if (a.debug_filt && valid) *reinterpret_cast<float4*>(a.filt + r * 192 + a.col0 + n0 + q * 4) = o;
          }
          group_sync(1 + g, F16_GROUP);    // half-tile complete
          aggregate_half<F>(a, rc, s_W, s_src, s_dst, s_runs, pass, s_carry + g * 128, s_carry + (1 - g) * 128, &bars[5 + g],
                            &bars[7 + g], &bars[5 + (1 - g)], &bars[7 + (1 - g)], prod_cnt, cons_cnt, gwarp, lane);
          group_sync(1 + g, F16_GROUP);    // half-tile consumed (s_W, and after the last pass s_src / s_dst / s_runs, may be rewritten)
        }
        if (rc.carry_out) ++prod_cnt;      // every warp of the group counts the same hand-offs
        if (rc.carry_in) ++cons_cnt;
      } else {
#pragma unroll
        for (int c = 0; c < HC / 32; ++c) {
          const int n0 = half * HC + c * 32;
          uint32_t v[32];
          tmem_ld32(trow + C16_D + n0, v);
          wait_ld();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.filt + r * 192 + a.col0 + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 o;
              o.x = fmaf(__uint_as_float(v[q * 4 + 0]), inv2, s_b2[n0 + q * 4 + 0]) * cw;
              o.y = fmaf(__uint_as_float(v[q * 4 + 1]), inv2, s_b2[n0 + q * 4 + 1]) * cw;
              o.z = fmaf(__uint_as_float(v[q * 4 + 2]), inv2, s_b2[n0 + q * 4 + 2]) * cw;
              o.w = fmaf(__uint_as_float(v[q * 4 + 3]), inv2, s_b2[n0 + q * 4 + 3]) * cw;
              __stcs(dst + q, o);
            }
          }
        }
      }
      // the next tile's operand stores target the A columns (dead since layer 2 completed); D is only overwritten after this
      // thread's next arrive on a_ready, which follows the wait::ld above in program order
    }
    if (amax > F16_RANGE) atomicOr(a.range_flag, 1);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int env_flag(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? (e[0] != '0') : dflt;
}
static int f16_scaled_default() { return env_flag("AGD_F16_LOSHIFT", 1); }
int f16_fuse_default() { return env_flag("AGD_F16_FUSE", 1); }

template <int F, bool FUSE>
static void launch_one(const TcF16Args& a, int grid, cudaStream_t s) {
  tc_filter16_kernel<F, FUSE><<<grid, F16_THREADS, TcF16Smem<F, FUSE>::bytes, s>>>(a);
}

void launch_filters_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& mw, int blk) {
  const BlkW& w = mw.blk[blk];
  static const int scaled = f16_scaled_default();
  const int fuse = c.f16_fuse;
  TcF16Args a{};
  a.n_rows_dev = b.counters;
  a.g2h = b.g2h;
  a.e_len = b.e_len;
  a.filt = b.filt;
  a.cutoff = c.cutoff;
  a.smooth = c.smooth;
  a.scaled = scaled;
  a.range_flag = b.counters + 4;
  a.debug_filt = c.f16_debug_filt;
  a.xcat = b.xcat;
  a.agg = b.agg;
  a.e_src = b.e_src;
  a.e_dst = b.e_dst;
  a.in_ptr = b.in_ptr;
  if (fuse) cudaMemsetAsync(b.agg, 0, sizeof(float) * 192 * (size_t)b.n_atoms, c.stream);   // atoms without in-edges
  int64_t pairs = (b.cap + 2 * TM - 1) / (2 * TM);
  const int grid = (int)(pairs < c.num_sms ? (pairs < 1 ? 1 : pairs) : c.num_sms);
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1a); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2a);
  a.f1b = w.f1ab; a.f2b = w.f2ab; a.dw = w.dw1; a.beta_ptr = w.sc + 0; a.wsc = w.hsc + 0; a.col0 = 0;
  if (fuse) launch_one<128, true>(a, grid, c.stream); else launch_one<128, false>(a, grid, c.stream);
  note_launch(c, fuse ? "schnet.cfconv128_f16" : "schnet.filter128_f16");
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1b); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2b);
  a.f1b = w.f1bb; a.f2b = w.f2bb; a.dw = w.dw2; a.beta_ptr = w.sc + 1; a.wsc = w.hsc + 2; a.col0 = 128;
  if (fuse) launch_one<64, true>(a, grid, c.stream); else launch_one<64, false>(a, grid, c.stream);
  note_launch(c, fuse ? "schnet.cfconv64_f16" : "schnet.filter64_f16");
}

int f16_lo_shift() { return f16_scaled_default() ? F16_LO_SHIFT : 0; }

void set_tc16_attributes() {
  cudaFuncSetAttribute(tc_filter16_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcF16Smem<128, false>::bytes);
  cudaFuncSetAttribute(tc_filter16_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcF16Smem<64, false>::bytes);
  cudaFuncSetAttribute(tc_filter16_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcF16Smem<128, true>::bytes);
  cudaFuncSetAttribute(tc_filter16_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcF16Smem<64, true>::bytes);
}

}  // namespace agd
