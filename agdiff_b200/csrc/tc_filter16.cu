// CFConv filter network on tcgen05 with fp16-split operands, TWO edge tiles in flight per SM (schnet.py:136-162, merged
// as in pack.py), UNFUSED: the filter tensor goes to HBM for cfconv_aggregate_kernel.  This is the A/B and test path of
// AGD_MODE_F16 ("f16_fuse" = 0); the default is tc_cfconv.cu, which runs both convs of a block and the aggregation in one
// warp-specialised launch and is checked against this file + the aggregate kernel bit for bit.
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
//   * epilogue 2: W = (D / s2 + b2) * cw -> filt.
// The envelope x distance-MLP weight cw_e of every edge and layer comes from edge_weight_kernel (once per evaluation).
#include <cuda_fp16.h>

#include <cstdlib>

#include "kernels.h"
#include "tc16_common.cuh"
#include "tc_filter16.cuh"

namespace agd {

template <int F>
struct TcF16Smem {
  static constexpr uint32_t W1_HALF = 128u * F * 2u, W2_HALF = static_cast<uint32_t>(F) * F * 2u;
  static constexpr size_t bytes = 1024 + 2 * W1_HALF + 2 * W2_HALF + (128 + 128) * sizeof(float) + 16 * sizeof(uint64_t) + 64;
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

template <int F>
__global__ void __launch_bounds__(F16_THREADS, 1) tc_filter16_kernel(const TcF16Args a) {
  using namespace tc;
  using SM = TcF16Smem<F>;
  constexpr uint32_t W1_HALF = SM::W1_HALF, W2_HALF = SM::W2_HALF;
  constexpr int HC = F / 2;   // output columns owned by one thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w1 = base;
  uint8_t* w2 = base + 2 * W1_HALF;
  float* s_b1 = reinterpret_cast<float*>(w2 + 2 * W2_HALF);
  float* s_b2 = s_b1 + 128;
  // [0] weights landed, [1+g] operand ready, [3+g] accumulator ready
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + 128);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = *a.n_rows_dev;

  // this CTA's contiguous row range
  const int64_t rows_per_cta = ((static_cast<int64_t>(n_rows) + TM - 1) / TM + gridDim.x - 1) / gridDim.x * TM;
  int64_t cta_begin = static_cast<int64_t>(blockIdx.x) * rows_per_cta, cta_end = cta_begin + rows_per_cta;
  if (cta_begin > n_rows) cta_begin = n_rows;
  if (cta_end > n_rows) cta_end = n_rows;
  const int cta_tiles = static_cast<int>((cta_end - cta_begin + TM - 1) / TM);   // <= 0: nothing to do

  if (warp == 0) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], F16_GROUP);
    mbar_init(&bars[2], F16_GROUP);
    mbar_init(&bars[3], 1);
    mbar_init(&bars[4], 1);
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
    const float inv1 = __ldg(a.wsc + 0) * (__ldg(a.beta_ptr) * 1.4426950408889634f), inv2 = __ldg(a.wsc + 1);
    const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    uint32_t dph = 0, aph = 0;
    const bool issuer = (gwarp == F16_GWARPS - 1) && lane == 0;
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
    // prefetched operand rows -> the slot's operand columns (called once layer 2 of the previous tile has completed)
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

    prefetch(g);
    stage_operand();
    float len_cur = pre_len;
    for (int j = g; j < cta_tiles; j += 2) {
      const int64_t row0 = cta_begin + static_cast<int64_t>(j) * TM;
      const int n_valid = (cta_end - row0 < TM) ? static_cast<int>(cta_end - row0) : TM;
      const int64_t r = row0 + my_row;
      const bool valid = my_row < n_valid;
      // ---- layer 1: the operand rows are already in the slot (stage_operand)
      wait_st();
      fence_before_sync();
      mbar_arrive(a_ready);
      if (issuer) {
        mbar_wait(a_ready, aph);
        aph ^= 1u;
        fence_after_sync();
        issue_3xf16<HID, F>(slot, smem_u32(w1), W1_HALF, scaled);
        mma_commit(d_ready);
      }
      const float cw = valid ? len_cur : 0.f;   // envelope * distance weight of this thread's row (edge_weight_kernel)
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
          const float t0 = ssp_log2(fmaf(__uint_as_float(v[2 * q]), inv1, s_b1[n0 + 2 * q]));
          const float t1 = ssp_log2(fmaf(__uint_as_float(v[2 * q + 1]), inv1, s_b1[n0 + 2 * q + 1]));
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
      stage_operand();   // next tile's operand rows (prefetched during layer 2) -> TMEM
      len_cur = pre_len;
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
      // the next tile's operand stores target the A columns (dead since layer 2 completed); D is only overwritten after this
      // thread's next arrive on a_ready, which follows the wait::ld above in program order
    }
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
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

template <int F>
static void launch_one(const TcF16Args& a, int grid, cudaStream_t s) {
  tc_filter16_kernel<F><<<grid, F16_THREADS, TcF16Smem<F>::bytes, s>>>(a);
}

// unfused path ("f16_fuse" = 0): filter tensor -> filt, aggregated by cfconv_aggregate_kernel (launch_aggregate)
void launch_filters_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& mw, int blk) {
  const BlkW& w = mw.blk[blk];
  static const int scaled = f16_scaled_default();
  TcF16Args a{};
  a.n_rows_dev = b.counters;
  a.g2h = b.g2h;
  a.filt = b.filt;
  a.scaled = scaled;
  a.range_flag = b.counters + 4;
  int64_t pairs = (b.cap + 2 * TM - 1) / (2 * TM);
  const int grid = (int)(pairs < c.num_sms ? (pairs < 1 ? 1 : pairs) : c.num_sms);
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1a); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2a);
  a.f1b = w.f1ab; a.f2b = w.f2ab; a.cw = b.cw_all + (size_t)(2 * blk) * (b.cap > 0 ? b.cap : 1); a.beta_ptr = w.sc + 0; a.wsc = w.hsc + 0; a.col0 = 0;
  launch_one<128>(a, grid, c.stream);
  note_launch(c, "schnet.filter128_f16");
  a.W1img = reinterpret_cast<const uint32_t*>(w.hF1b); a.W2img = reinterpret_cast<const uint32_t*>(w.hF2b);
  a.f1b = w.f1bb; a.f2b = w.f2bb; a.cw = b.cw_all + (size_t)(2 * blk + 1) * (b.cap > 0 ? b.cap : 1); a.beta_ptr = w.sc + 1; a.wsc = w.hsc + 2; a.col0 = 128;
  launch_one<64>(a, grid, c.stream);
  note_launch(c, "schnet.filter64_f16");
}

int f16_lo_shift() { return f16_scaled_default() ? F16_LO_SHIFT : 0; }

void set_tc16_attributes() {
  cudaFuncSetAttribute(tc_filter16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcF16Smem<128>::bytes);
  cudaFuncSetAttribute(tc_filter16_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcF16Smem<64>::bytes);
}

}  // namespace agd
