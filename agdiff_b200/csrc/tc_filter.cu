// CFConv filter-generating network on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   W_e = ( F2 . SSP_beta( F1 . g2_e + b1 ) + b2 ) * cw_e          (schnet.py:136-151, merged as in pack.py)
//
// fp32 fidelity on tensor cores: plain TF32 fails the 1e-4 parity bar, so every
// GEMM is 3xTF32:  A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi  with x_hi = x & 0xffffe000 (what the tensor
// core reads anyway) and x_lo = x - x_hi (exact in fp32), accumulated in fp32 in TMEM.
//
// Data flow of one 128-edge tile (one CTA per SM, 256 threads, 512 TMEM columns):
//   * the activation operand lives in TENSOR MEMORY (tcgen05.mma with A from TMEM): thread (warp w, lane l) owns
//     TMEM lane 32*(w%4)+l = tile row, warps 0-3 / 4-7 own the two column halves.  g2 rows are split in registers
//     and tcgen05.st'd as A_hi (cols 128..255) / A_lo (cols 256..383); the layer-1 accumulator D (cols 0..127)
//     is read back with tcgen05.ld, bias + ShiftedSoftplus applied, split again and stored over A for layer 2.
//     The MLP chain never touches shared or global memory.
//   * the weight operand is a pre-swizzled K-major SWIZZLE_128B image (hi | lo) built on the host; one thread
//     streams it from L2 with cp.async.bulk + mbarrier complete_tx into a single 128 KB buffer: layer-2 weights
//     are fetched while the layer-1 epilogue runs, the next tile's layer-1 weights during the layer-2 epilogue.
//   * one elected thread issues the 3 x K/8 tcgen05.mma (M=128, N=F, K=8) per layer and tcgen05.commit's to an
//     mbarrier that the 256 epilogue threads wait on.
#include "common.cuh"
#include "kernels.h"

namespace agd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion reported to an mbarrier (async proxy, like TMA)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 B, 8-row groups
// 1024 B apart (SBO), LBO unused (=1), descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, both K-major, M=128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr uint32_t TF32_MASK = 0xFFFFE000u;
constexpr int COL_D = 0, COL_AHI = 128, COL_ALO = 256, TMEM_COLS = 512;

}  // namespace tc

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
};

constexpr size_t TC_FILT_SMEM = 1024 /*align slack*/ + 131072 /*weights hi|lo*/ + 128 * sizeof(float) + 64;

template <int F>
__global__ void __launch_bounds__(256, 1) tc_filter_kernel(const TcFiltArgs a) {
  using namespace tc;
  constexpr uint32_t W1_HALF = (HID / 32) * F * 128;   // bytes of one (hi or lo) image of F1: K=128
  constexpr uint32_t W2_HALF = (F / 32) * F * 128;     // K=F
  constexpr int HALF_COLS = F / 2;                     // columns owned by one warp group
  constexpr int CHUNKS = HALF_COLS / 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* wbuf = base;                                             // 128 KB, 1024-aligned
  float* s_cw = reinterpret_cast<float*>(base + 131072);            // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_cw + 128);         // [0]=weights landed, [1]=mma done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
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
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint32_t trow = tmem + (static_cast<uint32_t>(quad * 32) << 16);   // this warp's lane quadrant
  uint32_t w_phase = 0, m_phase = 0;
  const float beta = __ldg(a.beta_ptr);

  auto load_weights = [&](const float* img, uint32_t half_bytes) {   // tid 0 only: hi then lo, 16 KB pieces
    mbar_expect_tx(&bars[0], 2 * half_bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img);
    for (uint32_t off = 0; off < 2 * half_bytes; off += 16384) bulk_g2s(wbuf + off, src + off, 16384, &bars[0]);
  };
  auto issue_layer = [&](int K, uint32_t half_bytes) {               // tid 0 only
    const uint32_t idesc = idesc_tf32(F);
    const uint32_t b_hi = smem_u32(wbuf), b_lo = smem_u32(wbuf) + half_bytes;
    for (int kb = 0; kb < K / 8; ++kb) {
      const uint32_t boff = static_cast<uint32_t>(kb >> 2) * (F * 128) + static_cast<uint32_t>(kb & 3) * 32;
      const uint64_t dh = smem_desc_sw128(b_hi + boff), dl = smem_desc_sw128(b_lo + boff);
      const uint32_t a_hi = tmem + COL_AHI + kb * 8, a_lo = tmem + COL_ALO + kb * 8;
      mma_tf32_ts(tmem + COL_D, a_hi, dh, idesc, kb > 0 ? 1u : 0u);
      mma_tf32_ts(tmem + COL_D, a_hi, dl, idesc, 1u);
      mma_tf32_ts(tmem + COL_D, a_lo, dh, idesc, 1u);
    }
    mma_commit(&bars[1]);
  };

  if (tid == 0 && blockIdx.x < n_tiles) load_weights(a.W1img, W1_HALF);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = static_cast<int64_t>(tile) * TM;
    const int64_t r = row0 + my_row;
    const bool valid = r < n_rows;
    // ---- stage A = g2 tile (hi/lo) into TMEM; per-edge envelope weight into smem
    if (half == 0) s_cw[my_row] = valid ? cfconv_edge_weight(a.e_len[r], a.dw, a.cutoff, a.smooth) : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {   // K = 128 input features: each warp group stores 64 columns
      uint32_t hi[16], lo[16];
      const float4* src = reinterpret_cast<const float4*>(a.g2 + r * HID + half * 64 + c * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t h = __float_as_uint(vv[j]) & TF32_MASK;
          hi[q * 4 + j] = h;
          lo[q * 4 + j] = __float_as_uint(vv[j] - __uint_as_float(h));
        }
      }
      tmem_st16(trow + COL_AHI + half * 64 + c * 16, hi);
      tmem_st16(trow + COL_ALO + half * 64 + c * 16, lo);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    // ---- layer 1 on the tensor core
    if (tid == 0) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      issue_layer(HID, W1_HALF);
    }
    w_phase ^= 1;
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
    if (tid == 0) load_weights(a.W2img, W2_HALF);   // layer-1 MMAs are complete: the weight buffer is free
    // ---- epilogue 1: t = SSP_beta(D + b1) -> A (hi/lo) for layer 2
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      uint32_t v[16], hi[16], lo[16];
      const int n0 = half * HALF_COLS + c * 16;
      tmem_ld16(trow + COL_D + n0, v);
      wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float t = ssp(__uint_as_float(v[j]) + __ldg(a.f1b + n0 + j), beta);
        const uint32_t h = __float_as_uint(t) & TF32_MASK;
        hi[j] = h;
        lo[j] = __float_as_uint(t - __uint_as_float(h));
      }
      tmem_st16(trow + COL_AHI + n0, hi);
      tmem_st16(trow + COL_ALO + n0, lo);
    }
    wait_st();
    fence_before_sync();
    __syncthreads();
    // ---- layer 2
    if (tid == 0) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      issue_layer(F, W2_HALF);
    }
    w_phase ^= 1;
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
    if (tid == 0 && tile + static_cast<int>(gridDim.x) < n_tiles) load_weights(a.W1img, W1_HALF);
    // ---- epilogue 2: W = (D + b2) * cw -> global filt[e][col0 + n]
    const float cw = s_cw[my_row];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      uint32_t v[16];
      const int n0 = half * HALF_COLS + c * 16;
      tmem_ld16(trow + COL_D + n0, v);
      wait_ld();
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(a.filt + r * 192 + a.col0 + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o;
          o.x = (__uint_as_float(v[q * 4 + 0]) + __ldg(a.f2b + n0 + q * 4 + 0)) * cw;
          o.y = (__uint_as_float(v[q * 4 + 1]) + __ldg(a.f2b + n0 + q * 4 + 1)) * cw;
          o.z = (__uint_as_float(v[q * 4 + 2]) + __ldg(a.f2b + n0 + q * 4 + 2)) * cw;
          o.w = (__uint_as_float(v[q * 4 + 3]) + __ldg(a.f2b + n0 + q * 4 + 3)) * cw;
          dst[q] = o;
        }
      }
    }
    fence_before_sync();
    __syncthreads();   // D and s_cw are free for the next tile
    fence_after_sync();
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
  int64_t tiles = (b.cap + TM - 1) / TM;
  const int grid = (int)(tiles < c.num_sms ? (tiles < 1 ? 1 : tiles) : c.num_sms);
  a.W1img = w.tF1a; a.W2img = w.tF2a; a.f1b = w.f1ab; a.f2b = w.f2ab; a.dw = w.dw1; a.beta_ptr = w.sc + 0; a.col0 = 0;
  tc_filter_kernel<128><<<grid, 256, TC_FILT_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.filter128_tc");
  a.W1img = w.tF1b; a.W2img = w.tF2b; a.f1b = w.f1bb; a.f2b = w.f2bb; a.dw = w.dw2; a.beta_ptr = w.sc + 1; a.col0 = 128;
  tc_filter_kernel<64><<<grid, 256, TC_FILT_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.filter64_tc");
}

void set_tc_attributes() {
  cudaFuncSetAttribute(tc_filter_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_FILT_SMEM);
  cudaFuncSetAttribute(tc_filter_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_FILT_SMEM);
}

}  // namespace agd
