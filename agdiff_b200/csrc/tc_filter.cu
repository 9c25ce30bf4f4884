// CFConv filter-generating network on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   W_e = ( F2 . SSP_beta( F1 . g2_e + b1 ) + b2 ) * cw_e          (schnet.py:136-151, merged as in pack.py)
//
// fp32 fidelity on tensor cores: plain TF32 fails the 1e-4 parity bar, so every
// GEMM is 3xTF32:  A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi  with x_hi = rna_tf32(x) and x_lo = rna_tf32(x - x_hi)
// (round-to-nearest on both parts: 2^-22 relative, unbiased), accumulated in fp32 in TMEM.
//
// Data flow of one 128-edge tile (one CTA per SM, 512 threads, 512 TMEM columns):
//   * the activation operand lives in TENSOR MEMORY (tcgen05.mma with A from TMEM): thread (warp w, lane l) owns
//     TMEM lane 32*(w%4)+l = tile row, the four warp groups own one quarter of the columns each.  g2 rows are split in registers
//     and tcgen05.st'd as A_hi (cols 128..255) / A_lo (cols 256..383); the layer-1 accumulator D (cols 0..127)
//     is read back with tcgen05.ld, bias + ShiftedSoftplus applied, split again and stored over A for layer 2.
//     The MLP chain never touches shared or global memory.
//   * the weight operand is a pre-swizzled K-major SWIZZLE_128B image (hi | lo) built on the host; one thread
//     streams it from L2 with cp.async.bulk + mbarrier complete_tx into a single 128 KB buffer: layer-2 weights
//     are fetched while the layer-1 epilogue runs, the next tile's layer-1 weights during the layer-2 epilogue.
//   * one elected thread issues the 3 x K/8 tcgen05.mma (M=128, N=F, K=8) per layer and tcgen05.commit's to an
//     mbarrier that the 512 epilogue threads wait on.
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace agd {

struct TcFiltArgs {
  const float* W1img;   // [hi | lo] swizzled K-major images of F1: each (128/32) x F rows x 128 B
  const float* W2img;   // [hi | lo] images of F2: (F/32) x F rows x 128 B
  const float *f1b, *f2b, *dw, *beta_ptr;
  const int* n_rows_dev;
  const float *g2, *e_len;
  float* filt;
  int col0;
  float cutoff;
  int smooth;
  int debug_nostream;   // timing experiment only: load the weights once, never re-stream (WRONG results)
};

constexpr int TC_THREADS = 512;
constexpr size_t TC_FILT_SMEM = 1024 /*align slack*/ + 131072 /*weights hi|lo*/ + (256 + 128 + 128 + 128) * sizeof(float) + 64;

template <int F>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_filter_kernel(const TcFiltArgs a) {
  using namespace tc;
  constexpr uint32_t W1_HALF = (HID / 32) * F * 128;   // bytes of one (hi or lo) image of F1: K=128
  constexpr uint32_t W2_HALF = (F / 32) * F * 128;     // K=F
  constexpr int PART_COLS = F / 4;                     // output columns owned by one of the 4 warp groups
  constexpr int CHUNKS = PART_COLS / 16;               // 2 (F=128) or 1 (F=64)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS, not generic LD)
  uint8_t* wbuf = base;                                             // 128 KB, 1024-aligned
  float* s_cw = reinterpret_cast<float*>(base + 131072);            // [2][128] envelope weight of each tile row (double-buffered)
  float* s_b1 = s_cw + 256;                                         // [F] layer-1 bias
  float* s_b2 = s_b1 + 128;                                         // [F] layer-2 bias
  float* s_dw = s_b2 + 128;                                         // [128] distance-weighting MLP
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dw + 128);         // [0]=weights landed, [1]=mma done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int my_row = quad * 32 + lane;
  const int n_rows = *a.n_rows_dev;
  const int n_tiles = (n_rows + TM - 1) / TM;

  if (warp == 0) {
    tmem_alloc(s_tmem, TMEM_COLS);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (tid < F) {
    s_b1[tid] = __ldg(a.f1b + tid);
    s_b2[tid] = __ldg(a.f2b + tid);
  }
  if (tid < 128) s_dw[tid] = __ldg(a.dw + tid);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);   // this warp's lane quadrant
  uint32_t w_phase = 0, m_phase = 0;
  const float beta = __ldg(a.beta_ptr);

  bool first_load = true;
  auto load_weights = [&](const float* img, uint32_t half_bytes) {   // tid 0 only: hi then lo, 16 KB pieces
    if (a.debug_nostream && !first_load) {   // keep the barrier protocol, move 16 bytes instead of 128 KB
      mbar_expect_tx(&bars[0], 16);
      bulk_g2s(wbuf, img, 16, &bars[0]);
      return;
    }
    first_load = false;
    mbar_expect_tx(&bars[0], 2 * half_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img);
    for (uint32_t off = 0; off < 2 * half_bytes; off += 16384) bulk_g2s(wbuf + off, src + off, 16384, &bars[0]);
  };
  // one accumulator (SPLIT=false): these outputs are scaled by the envelope (<= 1) and summed over ~33 edges downstream,
  // the measured error stays at the fp32 noise level
  auto issue_layer1 = [&]() {   // tid 0 only
    issue_3xtf32<HID, F, false>(tmem, smem_u32(wbuf), false);
    mma_commit(&bars[1]);
  };
  auto issue_layer2 = [&]() {
    issue_3xtf32<F, F, false>(tmem, smem_u32(wbuf), false);
    mma_commit(&bars[1]);
  };
  // this thread's slice of a g2 row: 32 input features = 8 x float4, kept in registers one tile ahead
  float4 pre[8];
  float pre_len = 0.f;
  auto prefetch = [&](int t) {
    const int64_t r = static_cast<int64_t>(t) * TM + my_row;
    if (t < n_tiles && r < n_rows) {
      const float4* src = reinterpret_cast<const float4*>(a.g2 + r * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) pre[q] = __ldg(src + q);
      pre_len = __ldg(a.e_len + r);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) pre[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      pre_len = -1.f;
    }
  };

  if (tid == 0 && blockIdx.x < n_tiles) load_weights(a.W1img, W1_HALF);
  prefetch(blockIdx.x);

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    float* cwbuf = s_cw + (it & 1) * 128;
    const int64_t row0 = static_cast<int64_t>(tile) * TM;
    const int64_t r = row0 + my_row;
    const bool valid = r < n_rows;
    // ---- stage A = g2 tile (hi/lo) into TMEM from the prefetched registers; envelope weight into smem
    if (part == 3) cwbuf[my_row] = valid ? cfconv_edge_weight_smem(pre_len, s_dw, a.cutoff, a.smooth) : 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = pre[c * 4 + q];
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          split_tf32(vv[j], hi[q * 4 + j], lo[q * 4 + j]);
        }
      }
      tmem_st16(trow + COL_AHI + part * 32 + c * 16, hi);
      tmem_st16(trow + COL_ALO + part * 32 + c * 16, lo);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    // ---- layer 1 on the tensor core
    if (tid == 0) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      issue_layer1();
    }
    w_phase ^= 1;
    prefetch(tile + static_cast<int>(gridDim.x));    // next tile's operand rows travel while the tensor core works
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
    if (tid == 0) load_weights(a.W2img, W2_HALF);   // layer-1 MMAs are complete: the weight buffer is free
    // ---- epilogue 1: t = SSP_beta(D + b1) -> A (hi/lo) for layer 2
    {
      float v[16 * CHUNKS];
      const int nb = part * PART_COLS;
      if constexpr (CHUNKS == 2) tmem_ld32_main(trow, nb, v); else tmem_ld16_main(trow, nb, v);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        uint32_t hi[16], lo[16];
        const int n0 = nb + c * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = ssp_fast(v[c * 16 + j] + s_b1[n0 + j], beta);
          split_tf32(t, hi[j], lo[j]);
        }
        tmem_st16(trow + COL_AHI + n0, hi);
        tmem_st16(trow + COL_ALO + n0, lo);
      }
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    // ---- layer 2
    if (tid == 0) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      issue_layer2();
    }
    w_phase ^= 1;
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
    if (tid == 0 && tile + static_cast<int>(gridDim.x) < n_tiles) load_weights(a.W1img, W1_HALF);
    // ---- epilogue 2: W = (D + b2) * cw -> global filt[e][col0 + n]
    const float cw = cwbuf[my_row];
    {
      float v[16 * CHUNKS];
      const int nb = part * PART_COLS;
      if constexpr (CHUNKS == 2) tmem_ld32_main(trow, nb, v); else tmem_ld16_main(trow, nb, v);
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(a.filt + r * 192 + a.col0 + nb);
#pragma unroll
        for (int q = 0; q < 4 * CHUNKS; ++q) {
          float4 o;
          o.x = (v[q * 4 + 0] + s_b2[nb + q * 4 + 0]) * cw;
          o.y = (v[q * 4 + 1] + s_b2[nb + q * 4 + 1]) * cw;
          o.z = (v[q * 4 + 2] + s_b2[nb + q * 4 + 2]) * cw;
          o.w = (v[q * 4 + 3] + s_b2[nb + q * 4 + 3]) * cw;
          __stcs(dst + q, o);
        }
      }
    }
    // no barrier here: the next tile's operand stores target the A columns (dead since layer 2 completed), its envelope
    // weights go to the other s_cw buffer, and D is only overwritten after the next tile's post-staging barrier
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

void launch_filters_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& mw, int blk) {
  const BlkW& w = mw.blk[blk];
  TcFiltArgs a{};
  a.n_rows_dev = b.counters;
  a.g2 = b.g2;
  a.e_len = b.e_len;
  a.filt = b.filt;
  a.cutoff = c.cutoff;
  a.smooth = c.smooth;
  { const char* e = getenv("AGD_TC_NOSTREAM"); a.debug_nostream = (e && e[0] == '1') ? 1 : 0; }
  int64_t tiles = (b.cap + TM - 1) / TM;
  const int grid = (int)(tiles < c.num_sms ? (tiles < 1 ? 1 : tiles) : c.num_sms);
  a.W1img = w.tF1a; a.W2img = w.tF2a; a.f1b = w.f1ab; a.f2b = w.f2ab; a.dw = w.dw1; a.beta_ptr = w.sc + 0; a.col0 = 0;
  tc_filter_kernel<128><<<grid, TC_THREADS, TC_FILT_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.filter128_tc");
  a.W1img = w.tF1b; a.W2img = w.tF2b; a.f1b = w.f1bb; a.f2b = w.f2bb; a.dw = w.dw2; a.beta_ptr = w.sc + 1; a.col0 = 128;
  tc_filter_kernel<64><<<grid, TC_THREADS, TC_FILT_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.filter64_tc");
}

void set_tc_attributes() {
  cudaFuncSetAttribute(tc_filter_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_FILT_SMEM);
  cudaFuncSetAttribute(tc_filter_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_FILT_SMEM);
}

}  // namespace agd
