// Shared pieces of the fp16-split ("3xFP16") two-slot tcgen05 kernels (tc_filter16.cu, tc_mlp16.cu): kind::f16 MMA issue with
// the scale-input-d fold of the cross terms, slot geometry, small PTX helpers.  See the header of tc_filter16.cu for the
// numerics and the pipeline.
#pragma once
#include <cuda_fp16.h>

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
// streaming 16-byte load that does not allocate in L1 (the g2h tile stream must not evict the x rows the aggregation gathers)
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
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
// one lane of a converged warp (the MMA issue below runs inside `if (elect_one())`: with a warp-uniform branch and compile-time
// TMEM addresses an MMA costs ~6 instructions; issued by a single thread of a diverged warp with run-time addresses it costs
// ~12 plus a uniform-branch loop per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t el;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(el));
  return el != 0;
}

// The same layers with the slot's TMEM columns as a template parameter (the kernel owns all 512 columns: the allocation starts
// at column 0, checked by the callers) - to be called by one elected lane.
template <int K, int N, uint32_t SLOT>
__device__ __forceinline__ void issue_3xf16_ct(uint32_t w_smem, uint32_t half_bytes, bool scaled) {
  constexpr uint32_t idesc = tc::idesc_f16(N);
  constexpr uint32_t D = SLOT + C16_D, ahi = SLOT + C16_AHI, alo = SLOT + C16_ALO;
  const uint64_t d_hi = tc::smem_desc_sw128(w_smem), d_lo = d_hi + (half_bytes >> 4);
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
template <int K, int N, uint32_t SLOT>
__device__ __forceinline__ void issue_3xf16_acc_ct(uint32_t w_smem, uint32_t half_bytes) {
  constexpr uint32_t idesc = tc::idesc_f16(N);
  constexpr uint32_t D = SLOT + C16_D, ahi = SLOT + C16_AHI, alo = SLOT + C16_ALO;
  const uint64_t d_hi = tc::smem_desc_sw128(w_smem), d_lo = d_hi + (half_bytes >> 4);
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    tc::mma_f16_ts(D, ahi + kb * 8, d_lo + boff16, idesc, 1u);
    tc::mma_f16_ts(D, alo + kb * 8, d_hi + boff16, idesc, 1u);
    tc::mma_f16_ts(D, ahi + kb * 8, d_hi + boff16, idesc, 1u);
  }
}

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

// D += A . W^T with UNSCALED lo parts (operand and weight image built with lo_shift = 0), accumulated onto a finished D: used
// where a second operand is added to an accumulator that already holds main-scale values (pair MLP, edge-feature half)
template <int K, int N>
__device__ __forceinline__ void issue_3xf16_acc(uint32_t slot, uint32_t w_smem, uint32_t half_bytes) {
  constexpr uint32_t idesc = tc::idesc_f16(N);
  const uint64_t d_hi = tc::smem_desc_sw128(w_smem), d_lo = tc::smem_desc_sw128(w_smem + half_bytes);
  const uint32_t D = slot + C16_D, ahi = slot + C16_AHI, alo = slot + C16_ALO;
#pragma unroll
  for (int kb = 0; kb < K / 16; ++kb) {
    const uint32_t boff16 = (static_cast<uint32_t>(kb >> 2) * (N * 128) + static_cast<uint32_t>(kb & 3) * 32) >> 4;
    tc::mma_f16_ts(D, ahi + kb * 8, d_lo + boff16, idesc, 1u);
    tc::mma_f16_ts(D, alo + kb * 8, d_hi + boff16, idesc, 1u);
    tc::mma_f16_ts(D, ahi + kb * 8, d_hi + boff16, idesc, 1u);
  }
}

}  // namespace agd
