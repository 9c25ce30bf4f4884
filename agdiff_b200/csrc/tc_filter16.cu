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
//   * warps 0-7 own slot 0, warps 8-15 slot 1; thread (quad q, lane l, half h) of a group owns tile row 32q+l and one half
//     of the columns.  The operand rows come pre-split from the encoder ("g2h", tc_common.cuh): they are prefetched into
//     registers during the previous tile's layer 2 and go straight into TMEM (tcgen05.st) - no ALU work;
//   * lane 0 of the group's last warp is its MMA issuer: once all 256 threads have arrived on the slot's "operand ready"
//     mbarrier it issues the layer's 3 x K/16 tcgen05.mma.kind::f16 (M=128, N=F, K=16, A from TMEM) and commits to the
//     slot's "accumulator ready" mbarrier.  While one group runs an epilogue the tensor core works on the other slot;
//   * epilogue 1: TMEM -> registers, bias + ShiftedSoftplus on the SFU, split, -> TMEM (operand of layer 2);
//   * epilogue 2 (FUSE): the filter tile is staged in shared memory as 64-column half-tiles and reduced per destination
//     (aggregate_half below); before that the accumulator is drained and the NEXT tile's layer 1 is issued, so that it runs
//     underneath the aggregation.  Without FUSE (A/B path, tests) the tile is written to `filt` for cfconv_aggregate_kernel.
// The envelope x distance-MLP weight cw_e of every edge and layer comes from edge_weight_kernel (once per evaluation).
#include <cuda_fp16.h>

#include <cstdlib>

#include "kernels.h"
#include "tc16_common.cuh"
#include "tc_filter16.cuh"

namespace agd {

template <int F, bool FUSE>
struct TcF16Smem {
  static constexpr uint32_t W1_HALF = 128u * F * 2u, W2_HALF = static_cast<uint32_t>(F) * F * 2u;
  static constexpr size_t fuse_bytes = FUSE ? (2 * TM * LDS_W + 2 * 128) * sizeof(float) + 6 * TM * sizeof(int) : 0;
  static constexpr size_t bytes = 1024 + 2 * W1_HALF + 2 * W2_HALF + (128 + 128) * sizeof(float) + fuse_bytes +
                                  16 * sizeof(uint64_t) + 64;
};

// CFConv edge weight lw(d) * C(d) (schnet.py:90-100,140-147) with the distance MLP staged in shared memory [w1 | b1 | w2 | b2],
// read as float4 (same arithmetic, in the same order, as cfconv_edge_weight_smem)
__device__ __forceinline__ float edge_weight_vec(float d, const float* dw, float cutoff, int smooth) {
  float z = dw[96];
  const float4* w1 = reinterpret_cast<const float4*>(dw);
  const float4* b1 = reinterpret_cast<const float4*>(dw + 32);
  const float4* w2 = reinterpret_cast<const float4*>(dw + 64);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 a4 = w1[q], b4 = b1[q], c4 = w2[q];
    z = fmaf(c4.x, relu_(fmaf(a4.x, d, b4.x)), z);
    z = fmaf(c4.y, relu_(fmaf(a4.y, d, b4.y)), z);
    z = fmaf(c4.z, relu_(fmaf(a4.z, d, b4.z)), z);
    z = fmaf(c4.w, relu_(fmaf(a4.w, d, b4.w)), z);
  }
  const float lw = sigmoidf_(z);
  float C;
  if (smooth) {
    C = 0.5f * (cosf(d * 3.14159265358979323846f / cutoff) + 1.0f);
    C = (d <= cutoff) ? C : 0.f;
  } else {
    const float t = d - cutoff;
    C = expf(-(t * t) / (2.0f * cutoff * cutoff));
  }
  C = (d <= cutoff && d >= 0.f) ? C : 0.f;
  return lw * C;
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
  const int colg = a.col0 + pass * 64 + 2 * lane;   // this lane's two columns in xcat / agg
  // descending, so that run 0 - the only one that may have to wait for the other group's carry - comes last
#pragma unroll 1
  for (int k = gwarp + ((rc.n_runs - 1 - gwarp) & ~(F16_GWARPS - 1)); k >= 0 && k < rc.n_runs; k -= F16_GWARPS) {
    const int s = s_runs[k];
    const int e = (k + 1 < rc.n_runs) ? s_runs[k + 1] : rc.n_valid;
    float2 acc = make_float2(0.f, 0.f);
    if (k == 0 && rc.carry_in) {
      tc::mbar_wait(full_other, cons_cnt & 1u);
      acc = *reinterpret_cast<const float2*>(carry_other + pass * 64 + 2 * lane);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty_other);
    }
    int row = s;
    for (; row + 16 <= e; row += 16) {
      float2 xv[16], wv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        xv[u] = __ldg(reinterpret_cast<const float2*>(a.xcat + (size_t)s_src[row + u] * 192 + colg));
        wv[u] = *reinterpret_cast<const float2*>(s_W + (row + u) * LDS_W + 2 * lane);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        acc.x = fmaf(xv[u].x, wv[u].x, acc.x);
        acc.y = fmaf(xv[u].y, wv[u].y, acc.y);
      }
    }
    if (row < e) {   // remainder (1..15 rows) as ONE predicated batch: a tail walked row by row costs an L2 round trip per row
      const int n = e - row;
      float2 xv[16], wv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        xv[u] = make_float2(0.f, 0.f);
        wv[u] = make_float2(0.f, 0.f);
        if (u < n) {
          xv[u] = __ldg(reinterpret_cast<const float2*>(a.xcat + (size_t)s_src[row + u] * 192 + colg));
          wv[u] = *reinterpret_cast<const float2*>(s_W + (row + u) * LDS_W + 2 * lane);
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float ax = fmaf(xv[u].x, wv[u].x, acc.x), ay = fmaf(xv[u].y, wv[u].y, acc.y);
        acc.x = (u < n) ? ax : acc.x;
        acc.y = (u < n) ? ay : acc.y;
      }
    }
    if (k == rc.n_runs - 1 && rc.carry_out) {
      tc::mbar_wait(empty_mine, (prod_cnt & 1u) ^ 1u);   // the previous carry of this group (same pass) was consumed
      *reinterpret_cast<float2*>(carry_mine + pass * 64 + 2 * lane) = acc;
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(full_mine);
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
  float* s_Wt = s_b2 + 128;                                        // FUSE: [2][128][LDS_W] filter half-tiles
  float* s_carry = s_Wt + (FUSE ? 2 * TM * LDS_W : 0);             // FUSE: [2][128] partial sums of runs cut by a tile boundary
  int* s_srcdst = reinterpret_cast<int*>(s_carry + (FUSE ? 256 : 0));   // FUSE: [2][3][128] src / dst / run starts of each slot's rows
  // [0] weights landed, [1+g] operand ready, [3+g] accumulator ready, [5+2g+pass] carry full, [9+2g+pass] carry empty
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_srcdst + (FUSE ? 6 * TM : 0));
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 14);

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
    for (int i = 3; i < 13; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (tid < F) {   // layer-1 bias pre-multiplied by beta * log2(e): the epilogue's first FFMA yields the exponent argument directly
    s_b1[tid] = __ldg(a.f1b + tid) * (__ldg(a.beta_ptr) * 1.4426950408889634f);
    s_b2[tid] = __ldg(a.f2b + tid);
  }
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
    float* s_W = s_Wt + g * TM * LDS_W;
    int* s_src = s_srcdst + g * 3 * TM;
    int* s_dst = s_src + TM;
    int* s_runs = s_dst + TM;
    const float inv1 = __ldg(a.wsc + 0) * (__ldg(a.beta_ptr) * 1.4426950408889634f), inv2 = __ldg(a.wsc + 1);
    const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    uint32_t dph = 0, aph = 0, prod_cnt = 0, cons_cnt = 0;
    const bool issuer = (gwarp == F16_GWARPS - 1) && lane == 0;   // the group's last warp rarely owns a run: its issue work stays off the aggregation's critical path
    const bool scaled = a.scaled != 0;
    if (issuer && g < cta_tiles) mbar_wait(&bars[0], 0);

    // this thread's slice of the pre-split g2 row: 32 hi words + 32 lo' words (K = 64*half .. 64*half+63), one tile ahead
    uint4 pre[16];
    float pre_len = 0.f;   // the row's edge weight cw, travels with the operand prefetch
    auto prefetch = [&](int j) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int64_t r = row0 + my_row;
      if (j < cta_tiles && r < cta_end) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          pre[q] = ldg_stream(a.g2h + g2h_index(r, 8 * half + q));
          pre[8 + q] = ldg_stream(a.g2h + g2h_index(r, 16 + 8 * half + q));
        }
        pre_len = __ldg(a.cw + r);
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) pre[q] = make_uint4(0u, 0u, 0u, 0u);
        pre_len = 0.f;
      }
    };

    // prefetched operand rows -> the slot's operand columns.  Called once layer 2 of the previous tile has completed (the A
    // columns are dead from then on), so the 64 prefetch registers are free again before the aggregation needs them.
    auto stage_operand = [&]() {
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
    };

    // phase timing (diagnostics only, compiled in with -DAGD_F16_TIMING: it costs ~20 registers): observers = first lane of
    // the group's warp 0 and warp 7
#ifdef AGD_F16_TIMING
    const bool obs = a.timing != nullptr && lane == 0 && (gwarp == 0 || gwarp == 7);
    long long t_prev = obs ? clock64() : 0;
    unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto tick = [&](int phase) {
      if (obs) {
        const long long t = clock64();
        t_acc[phase] += static_cast<unsigned long long>(t - t_prev);
        t_prev = t;
      }
    };
#else
    auto tick = [](int) {};
#endif

    prefetch(g);
    stage_operand();
    float len_cur = pre_len;
    for (int j = g; j < cta_tiles; j += 2) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      const int64_t r = row0 + my_row;
      const bool valid = my_row < n_valid;
      // ---- the operand rows are already in the slot (stage_operand); endpoints of the rows -> smem
      if (FUSE && half == 0) {
        s_src[my_row] = valid ? __ldg(a.e_src + r) : 0;
        s_dst[my_row] = valid ? __ldg(a.e_dst + r) : -1;
      }
      // layer 1 of this tile: with FUSE only the CTA's first tile of each slot is issued here - later ones are issued during the
      // previous tile's aggregation (below), which hides the whole MMA window
      const bool issue_here = !FUSE || j == g;
      if (issue_here) {
        wait_st();
        fence_before_sync();
      }
      group_sync(1 + g, F16_GROUP);   // publishes s_src / s_dst
      tick(0);
      if (issue_here) {
        mbar_arrive(a_ready);
        if (issuer) {
          mbar_wait(a_ready, aph);
          aph ^= 1u;
          fence_after_sync();
          issue_3xf16<HID, F>(slot, smem_u32(w1), W1_HALF, scaled);
          mma_commit(d_ready);
        }
      }
      const float cw = valid ? len_cur : 0.f;   // envelope * distance weight of this thread's row (edge_weight_kernel)
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
          const int start_rank = below + __popc(m & ((1u << lane) - 1u));
          if (gwarp == 0 && start) s_runs[start_rank] = row;   // published by the next group_sync
          // the x rows this warp will gather (runs gwarp, gwarp + 8, ...) -> L1 while the tensor core runs layer 1
          const int run = below + __popc(m & (0xffffffffu >> (31 - lane))) - 1;
          if (row < n_valid && (run & (F16_GWARPS - 1)) == gwarp) {
            const float* px = a.xcat + (size_t)s_src[row] * 192 + a.col0;
#pragma unroll
            for (int l = 0; l < F / 32; ++l) prefetch_l1(px + 32 * l);
          }
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
      tick(1);
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) {
        const int n0 = half * HC + c * 32;
        uint32_t v[32];
        tmem_ld32(trow + C16_D + n0, v);
        wait_ld();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float t0 = ssp_log2(fmaf(__uint_as_float(v[2 * q]), inv1, s_b1[n0 + 2 * q]));
          const float t1 = ssp_log2(fmaf(__uint_as_float(v[2 * q + 1]), inv1, s_b1[n0 + 2 * q + 1]));
          split2_f16(t0, t1, lo_scale, hi[q], lo[q], amax);
        }
        tmem_st16(trow + C16_AHI + (n0 >> 1), hi);
        tmem_st16(trow + C16_ALO + (n0 >> 1), lo);
      }
      wait_st();
      fence_before_sync();
      tick(2);
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
      tick(3);
      if (!FUSE) {
        stage_operand();   // next tile's operand rows (prefetched during layer 2) -> TMEM
        len_cur = pre_len;
      }
      if (FUSE) {
        // the filter tile is staged as 64-column half-tiles: pass p holds filter columns [64p, 64p+64)
        auto stage_half = [&](const uint32_t (&v)[32], int n0) {
          float4* dstW = reinterpret_cast<float4*>(s_W + my_row * LDS_W + half * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int n = n0 + q * 4;
            float4 o;
            o.x = fmaf(__uint_as_float(v[q * 4 + 0]), inv2, s_b2[n + 0]) * cw;
            o.y = fmaf(__uint_as_float(v[q * 4 + 1]), inv2, s_b2[n + 1]) * cw;
            o.z = fmaf(__uint_as_float(v[q * 4 + 2]), inv2, s_b2[n + 2]) * cw;
            o.w = fmaf(__uint_as_float(v[q * 4 + 3]), inv2, s_b2[n + 3]) * cw;
            dstW[q] = o;
            if (a.debug_filt && valid) *reinterpret_cast<float4*>(a.filt + r * 192 + a.col0 + n) = o;
          }
        };
        auto aggregate = [&](int pass) {
          aggregate_half<F>(a, rc, s_W, s_src, s_dst, s_runs, pass, s_carry + g * 128, s_carry + (1 - g) * 128,
                            &bars[5 + 2 * g + pass], &bars[9 + 2 * g + pass], &bars[5 + 2 * (1 - g) + pass],
                            &bars[9 + 2 * (1 - g) + pass], prod_cnt, cons_cnt, gwarp, lane);
        };
        {
          uint32_t v[32];
          tmem_ld32(trow + C16_D + half * 32, v);
          wait_ld();
          stage_half(v, half * 32);
        }
        stage_operand();   // next tile's operand rows (prefetched during layer 2) -> TMEM; frees the 64 prefetch registers
        len_cur = pre_len;
        uint32_t v1[32];   // F = 128: the second half-tile leaves TMEM now, so that the accumulator is free for the next tile
        if (F == 128) {
          tmem_ld32(trow + C16_D + 64 + half * 32, v1);
          wait_ld();
        }
        // D is fully read and the next operand is staged: the next tile's layer 1 runs underneath this tile's aggregation
        const bool next = j + 2 < cta_tiles;
        if (next) {
          wait_st();
          fence_before_sync();
          mbar_arrive(a_ready);
        }
        group_sync(1 + g, F16_GROUP);    // half-tile complete (and every thread of the group has arrived on a_ready)
        tick(4);
        if (next && issuer) {
          mbar_wait(a_ready, aph);
          aph ^= 1u;
          fence_after_sync();
          issue_3xf16<HID, F>(slot, smem_u32(w1), W1_HALF, scaled);
          mma_commit(d_ready);
        }
        aggregate(0);
        tick(5);
        group_sync(1 + g, F16_GROUP);    // half-tile consumed
        tick(6);
        if (F == 128) {
          stage_half(v1, 64 + half * 32);
          group_sync(1 + g, F16_GROUP);
          tick(4);
          aggregate(1);
          tick(5);
          group_sync(1 + g, F16_GROUP);  // s_W, s_src / s_dst / s_runs may be rewritten
          tick(6);
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
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
#ifdef AGD_F16_TIMING
    if (obs)
      for (int i = 0; i < 8; ++i) atomicAdd(a.timing + (g * 2 + (gwarp == 7 ? 1 : 0)) * 8 + i, t_acc[i]);
#endif
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- envelope * learnable distance weight of every edge for all CFConv layers of the step (schnet.py:90-100,140-147), once
// per network evaluation instead of inside each of the 12 filter launches (where its ~350 dependent instructions per row
// sat on the tile pipeline's critical path).  cw[net][e], net = 2 * block + conv.
struct EdgeWArgs {
  const float* dw[2 * MAX_BLOCKS];
  int n_nets;
  const int* n_rows_dev;
  const float* e_len;
  float* cw;
  int64_t stride;
  float cutoff;
  int smooth;
};
__global__ void __launch_bounds__(256) edge_weight_kernel(const EdgeWArgs a) {
  __shared__ __align__(16) float s_dw[2 * MAX_BLOCKS][128];
  for (int i = threadIdx.x; i < a.n_nets * 128; i += blockDim.x) s_dw[i >> 7][i & 127] = __ldg(a.dw[i >> 7] + (i & 127));
  __syncthreads();
  const int n = *a.n_rows_dev;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float d = __ldg(a.e_len + e);
    for (int k = 0; k < a.n_nets; ++k) a.cw[k * a.stride + e] = edge_weight_vec(d, s_dw[k], a.cutoff, a.smooth);
  }
}

void launch_edge_weights_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& mw) {
  EdgeWArgs a{};
  a.n_nets = 2 * c.num_convs;
  for (int k = 0; k < c.num_convs; ++k) {
    a.dw[2 * k] = mw.blk[k].dw1;
    a.dw[2 * k + 1] = mw.blk[k].dw2;
  }
  a.n_rows_dev = b.counters;
  a.e_len = b.e_len;
  a.cw = b.cw_all;
  a.stride = b.cap > 0 ? b.cap : 1;
  a.cutoff = c.cutoff;
  a.smooth = c.smooth;
  int64_t blocks = (b.cap + 255) / 256;
  const int grid = (int)(blocks < 1 ? 1 : (blocks > 8 * c.num_sms ? 8 * c.num_sms : blocks));
  edge_weight_kernel<<<grid, 256, 0, c.stream>>>(a);
  note_launch(c, "schnet.edge_weights");
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
  a.filt = b.filt;
  a.scaled = scaled;
  a.range_flag = b.counters + 4;
  a.debug_filt = c.f16_debug_filt;
  a.timing = c.f16_timing;
  a.xcat = b.xcat;
  a.agg = b.agg;
  a.e_src = b.e_src;
  a.e_dst = b.e_dst;
  a.in_ptr = b.in_ptr;
  // (atoms without in-edges: their agg rows stay unwritten, tc_node_kernel reads them as zero via in_ptr)
  int64_t pairs = (b.cap + 2 * TM - 1) / (2 * TM);
  const int grid = (int)(pairs < c.num_sms ? (pairs < 1 ? 1 : pairs) : c.num_sms);
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1a); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2a);
  a.f1b = w.f1ab; a.f2b = w.f2ab; a.cw = b.cw_all + (size_t)(2 * blk) * (b.cap > 0 ? b.cap : 1); a.beta_ptr = w.sc + 0; a.wsc = w.hsc + 0; a.col0 = 0;
  if (fuse) launch_one<128, true>(a, grid, c.stream); else launch_one<128, false>(a, grid, c.stream);
  note_launch(c, fuse ? "schnet.cfconv128_f16" : "schnet.filter128_f16");
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1b); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2b);
  a.f1b = w.f1bb; a.f2b = w.f2bb; a.cw = b.cw_all + (size_t)(2 * blk + 1) * (b.cap > 0 ? b.cap : 1); a.beta_ptr = w.sc + 1; a.wsc = w.hsc + 2; a.col0 = 128;
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
