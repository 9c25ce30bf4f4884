// tcgen05 / TMEM / mbarrier / bulk-copy primitives shared by the tensor-core kernels (sm_100a inline PTX;
// spellings follow the CUTLASS/CuTe sm100 headers, descriptor bit layouts cute::UMMA::{SmemDescriptor,InstrDescriptor}).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

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
// TMEM columns: main accumulator (hi.hi products), activation operand hi / lo, cross-term accumulator (hi.lo + lo.hi).
// The tensor core accumulates fp32 with truncation; keeping the 2^-11-smaller cross terms out of the main accumulator
// cuts its truncating steps 3x (they are summed in registers by the epilogue).
constexpr int COL_D = 0, COL_AHI = 128, COL_ALO = 256, COL_D2 = 384, TMEM_COLS = 512;

}  // namespace tc


// 1 / (1 + e^-x) on the SFU: ex2.approx + rcp.approx (relative error ~3e-7; e^-x -> inf gives exactly 0, NaN propagates)
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// softplus(beta*x) - ln2 on the SFU: ex2.approx / lg2.approx (absolute error ~4e-7; the parity bar is 1e-4 and the
// value feeds a 128-term fp32 dot product).  Matches F.softplus' threshold-20 branch.
__device__ __forceinline__ float ssp_fast(float x, float beta) {
  const float y = beta * x;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * 1.4426950408889634f));
  float sp;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(sp) : "f"(1.0f + e));
  sp = fmaf(sp, LN2F, -LN2F);
  return (y > 20.0f) ? (y - LN2F) : sp;
}

// accumulator chunk = main + cross terms
__device__ __forceinline__ void tmem_ld16_acc(uint32_t trow, int n0, float (&out)[16]) {
  uint32_t v[16], w[16];
  tc::tmem_ld16(trow + tc::COL_D + n0, v);
  tc::tmem_ld16(trow + tc::COL_D2 + n0, w);
  tc::wait_ld();
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
}

// single accumulator variant (kernels whose outputs are averaged downstream: filter nets, edge encoder)
__device__ __forceinline__ void tmem_ld16_main(uint32_t trow, int n0, float (&out)[16]) {
  uint32_t v[16];
  tc::tmem_ld16(trow + tc::COL_D + n0, v);
  tc::wait_ld();
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(v[j]);
}
__device__ __forceinline__ void tmem_ld32_main(uint32_t trow, int n0, float (&out)[32]) {
  uint32_t v[16], w[16];
  tc::tmem_ld16(trow + tc::COL_D + n0, v);
  tc::tmem_ld16(trow + tc::COL_D + n0 + 16, w);
  tc::wait_ld();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    out[j] = __uint_as_float(v[j]);
    out[16 + j] = __uint_as_float(w[j]);
  }
}

// exact-erf GELU (nn.GELU() default, edge.py:58,86) through the Gaussian tail: Phi(x) = 1 - q for x >= 0, q for x < 0, with
// q = 0.5 * poly(t) * exp(-x^2/2), t = 1/(1 + p|x|/sqrt2) (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7).  ~16 instructions
// (one rcp.approx, one ex2.approx) instead of ~40 for erff; measured max abs error 4.2e-7 over [-12, 12] - tighter than torch's
// own fp32 gelu (1.2e-6) because the negative side never forms 1 + erf(.) by cancellation.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, fabsf(z), 1.0f)));
  float p = 1.061405429f;
  p = fmaf(p, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-(z * z) * 1.4426950408889634f));
  const float q = 0.5f * p * e;
  return x * ((x >= 0.f) ? (1.0f - q) : q);
}

// The whole 3xTF32 issue sequence of one layer, D[128 x N] (+)= A[128 x K] . W^T, for ONE thread.  K and N are compile-time
// so the K/8 x 3 instructions are straight-line code: descriptors are the layer's base descriptor plus a constant (the
// 14-bit address field cannot carry: shared addresses are < 256 KB), instead of being rebuilt in a serial loop that is
// itself as long as the tensor core needs to execute the MMAs.  SPLIT: cross terms go to the second accumulator.
template <int K, int N, bool SPLIT>
__device__ __forceinline__ void issue_3xtf32(uint32_t tmem, uint32_t b_smem, bool accumulate) {
  constexpr uint32_t idesc = tc::idesc_tf32(N);
  constexpr uint32_t half_bytes = static_cast<uint32_t>(K) * N * 4;
  const uint64_t d_hi = tc::smem_desc_sw128(b_smem), d_lo = tc::smem_desc_sw128(b_smem + half_bytes);
  const uint32_t dmain = tmem + tc::COL_D, dx = tmem + (SPLIT ? tc::COL_D2 : tc::COL_D);
#pragma unroll
  for (int kb = 0; kb < K / 8; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    const uint64_t dh = d_hi + boff16, dl = d_lo + boff16;
    const uint32_t a_hi = tmem + tc::COL_AHI + kb * 8, a_lo = tmem + tc::COL_ALO + kb * 8;
    const uint32_t acc = (accumulate || kb > 0) ? 1u : 0u;
    tc::mma_tf32_ts(dmain, a_hi, dh, idesc, acc);
    tc::mma_tf32_ts(dx, a_hi, dl, idesc, SPLIT ? acc : 1u);
    tc::mma_tf32_ts(dx, a_lo, dh, idesc, 1u);
  }
}

// The same layer with the TMEM base as a compile-time 0 (a kernel that owns all 512 columns), to be called by ONE elected lane
// of a converged warp (tc16_common.cuh: elect_one): half the issue instructions of the single-diverged-thread form above.
template <int K, int N, bool SPLIT>
__device__ __forceinline__ void issue_3xtf32_ct(uint32_t b_smem, bool accumulate) {
  constexpr uint32_t idesc = tc::idesc_tf32(N);
  constexpr uint32_t half_bytes = static_cast<uint32_t>(K) * N * 4;
  const uint64_t d_hi = tc::smem_desc_sw128(b_smem), d_lo = d_hi + (half_bytes >> 4);
  constexpr uint32_t dmain = tc::COL_D, dx = SPLIT ? tc::COL_D2 : tc::COL_D;
#pragma unroll
  for (int kb = 0; kb < K / 8; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    const uint32_t acc = (accumulate || kb > 0) ? 1u : 0u;
    tc::mma_tf32_ts(dmain, tc::COL_AHI + kb * 8, d_hi + boff16, idesc, acc);
    tc::mma_tf32_ts(dx, tc::COL_AHI + kb * 8, d_lo + boff16, idesc, SPLIT ? acc : 1u);
    tc::mma_tf32_ts(dx, tc::COL_ALO + kb * 8, d_hi + boff16, idesc, 1u);
  }
}

// split an fp32 value into two TF32 values with round-to-nearest: hi = rna(v), lo = rna(v - hi) (v - hi is exact).
// |v - (hi + lo)| <= 2^-22 |v| and unbiased; feeding raw fp32 bits instead would let the tensor core TRUNCATE
// (2^-20, biased), which the far-geometry parity case does not tolerate.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float rem = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}

// ---- fp16 split ("3xFP16", tc_filter16.cu): two fp32 values -> packed fp16 hi pair and packed fp16 lo' pair
// (low half = first value), hi = rn_f16(x), lo' = rn_f16((x - hi) * lo_scale); amax tracks max|x| for the range check.
constexpr int F16_LO_SHIFT = 11;        // S: lo' = (x - hi) * 2^S (pack.umma_image_f16 builds the weight images with it)
constexpr float F16_RANGE = 65000.f;
__device__ __forceinline__ bool f16_out_of_range(__half2 amax) {
  const float2 f = __half22float2(amax);
  return fmaxf(f.x, f.y) > F16_RANGE;
}
__device__ __forceinline__ void split2_f16(float x0, float x1, float lo_scale, uint32_t& hi, uint32_t& lo, __half2& amax) {
  const __half2 h = __floats2half2_rn(x0, x1);
  amax = __hmax2(amax, __habs2(h));            // inf when |x| is beyond the fp16 range; NaN operands are ignored
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * lo_scale, (x1 - hf.y) * lo_scale);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// The pre-split copy of the encoder state g2 that the fp16 filter kernels read ("g2h"): per 128-edge block
// [word quad w4 (32)][row (128)][4 words], words 0..63 of a row = packed hi pairs of features (2w, 2w+1), words 64..127 the
// lo' pairs.  Producer and consumers are thread-per-row, so every warp access is 512 contiguous bytes.
__device__ __forceinline__ size_t g2h_index(int64_t row, int w4) {   // in uint4 units
  return (static_cast<size_t>(row >> 7) * 32 + w4) * 128 + static_cast<size_t>(row & 127);
}

}  // namespace agd
