// EXPERIMENTAL (opt-in: agd_set_option "f16_ws" / env AGD_F16_WS=1; the default CFConv kernel is tc_filter16.cu).  Written at the
// end of round 1 with the last GPU seconds of the round: FUNCTIONALLY validated - with AGD_F16_WS=1 the bit-for-bit tests against the
// stand-alone aggregate kernel pass (test_f16_fused_aggregation_bitwise_equals_unfused x3, test_supplied_complete_graph_long_runs) -
// but NOT yet timed or profiled, hence not the default.
//
// Warp-specialised CFConv kernel of the fp16-split family.  Same math, same tiles, same summation order as
// tc_filter16_kernel<F, fused> - only the division of labour changes.  There, each 8-warp group owns a tile end to end, so the
// group that is aggregating (L2-latency-bound gathers) cannot run its next epilogue, and ncu shows no pipe above 40 %: the kernel
// is bound by how little independent work 16 warps expose.  Here:
//
//   * warps 0-7 ("epilogue group") run, for EVERY tile of the CTA and alternating between the two TMEM slots: operand staging,
//     epilogue 1 (SSP, split, back into TMEM), the drain of the layer-2 accumulator into shared-memory half-tiles, and all MMA
//     issue.  Their schedule per tile j is  P1(j): wait layer 1 -> epilogue 1 -> issue layer 2 -> prefetch operand j+1;
//     P2(j-1): wait layer 2 of the OTHER slot -> drain -> stage operand j+1 there -> issue its layer 1 - so the tensor core
//     always has one slot's layer in flight while the group computes on the other;
//   * warps 8-15 ("aggregation group") consume the half-tiles through a two-deep ring (full/empty mbarriers) and reduce them
//     per destination run with one fmaf per edge in CSC order.  All tiles of the CTA pass through the same 8 warps in order, so
//     the partial sum of a run cut by a tile boundary is carried in shared memory between consecutive tiles of the group
//     (double-buffered by tile parity) - no cross-group hand-off at all.
#include "kernels.h"
#include "tc_filter16.cuh"

namespace agd {

using namespace tc;

constexpr int WS_THREADS = 512;
constexpr int WS_GROUP = 256;

template <int F>
struct TcWsSmem {
  static constexpr uint32_t W1_HALF = 128u * F * 2u, W2_HALF = static_cast<uint32_t>(F) * F * 2u;
  static constexpr size_t bytes = 1024 + 2 * W1_HALF + 2 * W2_HALF + (128 + 128) * sizeof(float) +
                                  (2 * TM * LDS_W + 2 * 128) * sizeof(float) + 3 * TM * sizeof(int) + 16 * sizeof(uint64_t) + 64;
};

// aggregation of one 64-column half-tile (pass `pass` of tile parity `tpar`); carry_prev / carry_cur: [2 passes][64] partial sums
// of the run cut by the previous / the next tile boundary
__device__ __forceinline__ void ws_aggregate(const TcF16Args& a, int n_runs, int n_valid, bool carry_in, bool carry_out, const float* s_W,
                                             const int* s_src, const int* s_dst, const int* s_runs, int pass, const float* carry_prev,
                                             float* carry_cur, int gwarp, int lane) {
  const int colg = a.col0 + pass * 64 + 2 * lane;
#pragma unroll 1
  for (int k = gwarp; k < n_runs; k += 8) {
    const int s = s_runs[k];
    const int e = (k + 1 < n_runs) ? s_runs[k + 1] : n_valid;
    float2 acc = make_float2(0.f, 0.f);
    if (k == 0 && carry_in) acc = *reinterpret_cast<const float2*>(carry_prev + pass * 64 + 2 * lane);
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
    if (row < e) {   // remainder as one predicated batch
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
    if (k == n_runs - 1 && carry_out) *reinterpret_cast<float2*>(carry_cur + pass * 64 + 2 * lane) = acc;
    else *reinterpret_cast<float2*>(a.agg + (size_t)s_dst[s] * 192 + colg) = acc;
  }
}

template <int F>
__global__ void __launch_bounds__(WS_THREADS, 1) tc_filter16_ws_kernel(const TcF16Args a) {
  using SM = TcWsSmem<F>;
  constexpr uint32_t W1_HALF = SM::W1_HALF, W2_HALF = SM::W2_HALF;
  constexpr int HC = F / 2;        // epilogue-1 columns per thread
  constexpr int PASSES = F / 64;   // half-tiles per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w1 = base;
  uint8_t* w2 = base + 2 * W1_HALF;
  float* s_b1 = reinterpret_cast<float*>(w2 + 2 * W2_HALF);
  float* s_b2 = s_b1 + 128;
  float* s_ring = s_b2 + 128;                           // [2][128][LDS_W] half-tile ring
  float* s_carry = s_ring + 2 * TM * LDS_W;             // [2 tile parities][2 passes][64]
  int* s_src = reinterpret_cast<int*>(s_carry + 256);   // [128] bookkeeping of the tile being aggregated
  int* s_dst = s_src + TM;
  int* s_runs = s_dst + TM;
  // [0] weights landed, [1+s] operand ready (256), [3+s] accumulator ready (1), [5+b] half-tile full (256), [7+b] half-tile empty (1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_runs + TM);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = *a.n_rows_dev;
  // this CTA's contiguous row range, snapped to run (destination) boundaries - identical to tc_filter16_kernel<F, true>
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
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], WS_GROUP);
    mbar_init(&bars[2], WS_GROUP);
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
    mbar_init(&bars[5], WS_GROUP);
    mbar_init(&bars[6], WS_GROUP);
    mbar_init(&bars[7], 1);
    mbar_init(&bars[8], 1);
    fence_barrier_init();
  }
  if (tid < F) {
    s_b1[tid] = __ldg(a.f1b + tid) * (__ldg(a.beta_ptr) * 1.4426950408889634f);
    s_b2[tid] = __ldg(a.f2b + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;

  if (tid == 0 && T > 0) {   // both layers' weight images, once
    mbar_expect_tx(&bars[0], 2 * W1_HALF + 2 * W2_HALF);
    const uint8_t* p1 = reinterpret_cast<const uint8_t*>(a.W1img);
    const uint8_t* p2 = reinterpret_cast<const uint8_t*>(a.W2img);
    for (uint32_t off = 0; off < 2 * W1_HALF; off += 16384) bulk_g2s(w1 + off, p1 + off, 16384, &bars[0]);
    for (uint32_t off = 0; off < 2 * W2_HALF; off += 16384) bulk_g2s(w2 + off, p2 + off, 16384, &bars[0]);
  }

  if (warp < 8) {
    // ================================================================== epilogue group
    const int quad = warp & 3, half = warp >> 2;
    const int my_row = quad * 32 + lane;
    const bool issuer = (tid == 0);
    const bool scaled = a.scaled != 0;
    const float inv1 = __ldg(a.wsc + 0) * (__ldg(a.beta_ptr) * 1.4426950408889634f), inv2 = __ldg(a.wsc + 1);
    const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    uint32_t dph = 0u, aph = 0u;   // phase bit of slot s = bit s (no dynamically indexed local arrays)
    uint32_t n_half = 0;          // half-tiles staged so far: ring position n_half & 1, use index n_half >> 1
    float cw_slot0 = 0.f, cw_slot1 = 0.f;   // edge weight of the row currently owned by slot 0 / 1
    if (issuer && T > 0) mbar_wait(&bars[0], 0);

    uint4 pre[16];
    float pre_cw = 0.f;
    auto load_operand = [&](int j) {   // this thread's 32 hi + 32 lo' words of tile j's row, and the row's edge weight
      const int64_t r = cta_begin + static_cast<int64_t>(j) * TM + my_row;
      if (j < T && r < cta_end) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          pre[q] = ldg_stream(a.g2h + g2h_index(r, 8 * half + q));
          pre[8 + q] = ldg_stream(a.g2h + g2h_index(r, 16 + 8 * half + q));
        }
        pre_cw = __ldg(a.cw + r);
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) pre[q] = make_uint4(0u, 0u, 0u, 0u);
        pre_cw = 0.f;
      }
    };
    auto stage_operand = [&](int s) {   // -> operand columns of slot s
      const uint32_t trow = tmem + static_cast<uint32_t>(s * SLOT_COLS) + (static_cast<uint32_t>(quad * 32) << 16);
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
    // publish this thread's operand stores of slot s and (issuer) run a layer on it
    auto run_layer = [&](int s, int layer) {
      wait_st();
      fence_before_sync();
      mbar_arrive(&bars[1 + s]);
      if (issuer) {
        mbar_wait(&bars[1 + s], (aph >> s) & 1u);
        fence_after_sync();
        const uint32_t slot = tmem + static_cast<uint32_t>(s * SLOT_COLS);
        if (layer == 0) issue_3xf16<HID, F>(slot, smem_u32(w1), W1_HALF, scaled);
        else issue_3xf16<F, F>(slot, smem_u32(w2), W2_HALF, scaled);
        mma_commit(&bars[3 + s]);
      }
      aph ^= (1u << s);
    };
    auto wait_acc = [&](int s) {
      mbar_wait(&bars[3 + s], (dph >> s) & 1u);
      dph ^= (1u << s);
      fence_after_sync();
    };
    // P2(t): drain the layer-2 accumulator of tile t into the half-tile ring; once it is drained, stage the prefetched operand
    // of tile t + 2 into the same slot and start its layer 1
    auto drain = [&](int t, bool has_next) {
      const int s = t & 1;
      const uint32_t trow = tmem + static_cast<uint32_t>(s * SLOT_COLS) + (static_cast<uint32_t>(quad * 32) << 16);
      const int64_t row0 = cta_begin + static_cast<int64_t>(t) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      const int64_t r = row0 + my_row;
      const bool valid = my_row < n_valid;
      const float cw = valid ? (s ? cw_slot1 : cw_slot0) : 0.f;
      wait_acc(s);
#pragma unroll
      for (int h = 0; h < PASSES; ++h) {
        const int n0 = h * 64 + half * 32;
        uint32_t v[32];
        tmem_ld32(trow + C16_D + n0, v);
        wait_ld();
        if (h == PASSES - 1 && has_next) {   // the accumulator is fully read: the slot can take its next tile
          stage_operand(s);
          if (s) cw_slot1 = pre_cw; else cw_slot0 = pre_cw;
          run_layer(s, 0);
        }
        const uint32_t b = n_half & 1u;
        mbar_wait(&bars[7 + b], ((n_half >> 1) & 1u) ^ 1u);   // the aggregation group released this ring slot
        float4* dstW = reinterpret_cast<float4*>(s_ring + b * TM * LDS_W + my_row * LDS_W + half * 32);
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
        mbar_arrive(&bars[5 + b]);   // release: this thread's part of the half-tile is in shared memory
        ++n_half;
      }
    };

    // prologue: layer 1 of the first two tiles
    for (int t = 0; t < 2 && t < T; ++t) {
      load_operand(t);
      stage_operand(t);
      if (t) cw_slot1 = pre_cw; else cw_slot0 = pre_cw;
      run_layer(t, 0);
    }
    for (int j = 0; j < T; ++j) {
      const int s = j & 1;
      const uint32_t trow = tmem + static_cast<uint32_t>(s * SLOT_COLS) + (static_cast<uint32_t>(quad * 32) << 16);
      // ---- P1(j): layer 1 done -> epilogue 1 -> layer 2
      wait_acc(s);
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
      run_layer(s, 1);
      // ---- P2(j-1) on the other slot, with the operand of tile j+1 (which goes into that slot) already travelling
      if (j >= 1) {
        load_operand(j + 1);
        drain(j - 1, j + 1 < T);
      }
    }
    if (T > 0) drain(T - 1, false);
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
  } else {
    // ================================================================== aggregation group
    const int gwarp = warp - 8, gtid = tid - WS_GROUP;
    uint32_t n_half = 0;
    for (int j = 0; j < T; ++j) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      if (gtid < TM) {
        const bool valid = gtid < n_valid;
        s_src[gtid] = valid ? __ldg(a.e_src + row0 + gtid) : 0;
        s_dst[gtid] = valid ? __ldg(a.e_dst + row0 + gtid) : -1;
      }
      group_sync(2, WS_GROUP);
      // run structure (every warp computes the same masks), run starts -> s_runs
      const int prev_dst = (j > 0) ? __ldg(a.e_dst + row0 - 1) : -2;
      int n_runs = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int row = w * 32 + lane;
        const int d = s_dst[row];
        const int dp = (row > 0) ? s_dst[row - 1] : -3;
        const bool start = row < n_valid && (row == 0 || d != dp);
        const uint32_t m = __ballot_sync(0xffffffffu, start);
        if (gwarp == 0 && start) s_runs[n_runs + __popc(m & ((1u << lane) - 1u))] = row;
        n_runs += __popc(m);
      }
      const bool carry_in = (s_dst[0] == prev_dst);
      const bool carry_out = (row0 + n_valid < cta_end) && (__ldg(a.e_dst + row0 + n_valid) == s_dst[n_valid - 1]);
      group_sync(2, WS_GROUP);   // s_runs visible
      const float* carry_prev = s_carry + ((j & 1) ^ 1) * 128;
      float* carry_cur = s_carry + (j & 1) * 128;
#pragma unroll 1
      for (int h = 0; h < PASSES; ++h) {
        const uint32_t b = n_half & 1u;
        mbar_wait(&bars[5 + b], (n_half >> 1) & 1u);   // half-tile staged by all 256 epilogue threads
        ws_aggregate(a, n_runs, n_valid, carry_in, carry_out, s_ring + b * TM * LDS_W, s_src, s_dst, s_runs, h, carry_prev, carry_cur,
                     gwarp, lane);
        group_sync(2, WS_GROUP);   // every warp is done with the ring slot (and, after the last pass, with the bookkeeping)
        if (gtid == 0) mbar_arrive(&bars[7 + b]);
        ++n_half;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

void launch_filters_f16_ws(const LaunchCtx& c, const BatchDev& b, const ModelW& mw, int blk) {
  const BlkW& w = mw.blk[blk];
  TcF16Args a{};
  a.n_rows_dev = b.counters;
  a.g2h = b.g2h;
  a.filt = b.filt;
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  a.debug_filt = c.f16_debug_filt;
  a.timing = nullptr;
  a.xcat = b.xcat;
  a.agg = b.agg;
  a.e_src = b.e_src;
  a.e_dst = b.e_dst;
  a.in_ptr = b.in_ptr;
  int64_t pairs = (b.cap + 2 * TM - 1) / (2 * TM);
  const int grid = (int)(pairs < c.num_sms ? (pairs < 1 ? 1 : pairs) : c.num_sms);
  const size_t stride = (size_t)(b.cap > 0 ? b.cap : 1);
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1a); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2a);
  a.f1b = w.f1ab; a.f2b = w.f2ab; a.cw = b.cw_all + (size_t)(2 * blk) * stride; a.beta_ptr = w.sc + 0; a.wsc = w.hsc + 0; a.col0 = 0;
  tc_filter16_ws_kernel<128><<<grid, WS_THREADS, TcWsSmem<128>::bytes, c.stream>>>(a);
  note_launch(c, "schnet.cfconv128_f16ws");
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1b); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2b);
  a.f1b = w.f1bb; a.f2b = w.f2bb; a.cw = b.cw_all + (size_t)(2 * blk + 1) * stride; a.beta_ptr = w.sc + 1; a.wsc = w.hsc + 2; a.col0 = 128;
  tc_filter16_ws_kernel<64><<<grid, WS_THREADS, TcWsSmem<64>::bytes, c.stream>>>(a);
  note_launch(c, "schnet.cfconv64_f16ws");
}

void set_tc16_ws_attributes() {
  cudaFuncSetAttribute(tc_filter16_ws_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcWsSmem<128>::bytes);
  cudaFuncSetAttribute(tc_filter16_ws_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcWsSmem<64>::bytes);
}

}  // namespace agd
