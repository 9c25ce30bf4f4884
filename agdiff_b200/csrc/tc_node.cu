// Node-side chain of a SchNet interaction block on tcgen05 (same 3xTF32 / TMEM-chained scheme as tc_mlp.cu):
//
//   t1 = SSP(BN(lin2_1(agg[:, :128])))            t2 = SSP(BN(lin2_2(agg[:, 128:])))           schnet.py:157-158,206-207
//   xc = lin([t1, t2])  (K = 256 as two K = 128 passes; the first partial waits in registers)     schnet.py:208
//   gate = sigmoid(a2 . relu(A1 xc + a1b) + a2b);  y = xc * gate                                 schnet.py:211-214
//   s = sigmoid(S2^T relu(S1^T y));  h += y * s                                                  schnet.py:230-234,278-280
//   next block: x = LeakyReLU_0.2(BN(lin1(h))) for conv1 (128) and conv2 (64)                    schnet.py:152-155
//
// One CTA per SM, 512 threads: thread (warp w, lane l) owns node row 32*(w%4)+l and the column quarter w/4; row-wise
// reductions (attention logit, the 128->8 squeeze) combine the four quarters through shared memory.
#include "kernels.h"
#include "tc_common.cuh"

namespace agd {

using namespace tc;

constexpr int TCN_THREADS = 512;

struct TcNodeArgs {
  BlkW w;                                   // block whose convs just aggregated (unused when first)
  const float *nL1a_img, *nL1b_img;         // NEXT block's lin1 images (nullptr after the last block)
  const float *nl1ab, *nl1bb;
  const float* emb;
  const int* atom_type;
  int n_nodes;
  int first;
  const float* agg;   // [N][192]
  const int* in_ptr;  // [N+1] in-edge segments: atoms without in-edges aggregate to zero (their agg rows are not written by the fused kernels)
  float* h;           // [N][128]
  float* xcat;        // [N][192]
};

constexpr size_t TC_NODE_SMEM = 1024 + 131072 + (128 * 3 + 64 * 2 + 128 + 64 + 1024 + 1024 + 512 + 4096) * sizeof(float) + 256;

__device__ __forceinline__ bool elect_one_tf32() {
  uint32_t el;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(el));
  return el != 0;
}

struct NodeCtx {   // TcCtx from tc_mlp.cu, repeated here to keep the translation units independent
  uint8_t* wbuf;
  uint64_t* bars;
  uint32_t tmem, trow;
  uint32_t w_phase, m_phase;
  int tid;
  __device__ __forceinline__ void stream(const float* img, uint32_t bytes) {
    mbar_expect_tx(&bars[0], bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img);
    for (uint32_t off = 0; off < bytes; off += 16384) bulk_g2s(wbuf + off, src + off, 16384, &bars[0]);
  }
  __device__ __forceinline__ void layer(int K, int N) {   // all threads; weights for this layer were streamed earlier
    wait_st();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      const uint32_t b_smem = smem_u32(wbuf);
      if (K == 128 && N == 128) issue_3xtf32<128, 128, true>(tmem, b_smem, false);
      else if (K == 64 && N == 128) issue_3xtf32<64, 128, true>(tmem, b_smem, false);
      else issue_3xtf32<128, 64, true>(tmem, b_smem, false);
      mma_commit(&bars[1]);
    }
    w_phase ^= 1;
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
  }
  // 128 x 128 layer issued by one elected lane of warp 0 with compile-time TMEM addresses (only valid when the kernel's TMEM
  // allocation starts at column 0 - the caller checks)
  __device__ __forceinline__ void layer128_elected() {
    wait_st();
    fence_before_sync();
    __syncthreads();
    if (tid < 32) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      if (elect_one_tf32()) {
        issue_3xtf32_ct<128, 128, true>(smem_u32(wbuf), false);
        mma_commit(&bars[1]);
      }
      __syncwarp();
    }
    w_phase ^= 1;
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
  }
};

__device__ __forceinline__ void st_split16(uint32_t trow, int k0, const float (&t)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split_tf32(t[j], hi[j], lo[j]);
  tmem_st16(trow + COL_AHI + k0, hi);
  tmem_st16(trow + COL_ALO + k0, lo);
}

__global__ void __launch_bounds__(TCN_THREADS, 1) tc_node_kernel(const TcNodeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS, not generic LD)
  float* s_l2ab = reinterpret_cast<float*>(base + 131072);
  float* s_l2bb = s_l2ab + 128;
  float* s_linb = s_l2bb + 128;
  float* s_a1b = s_linb + 128;    // [64]
  float* s_a2w = s_a1b + 64;      // [64]
  float* s_l1ab = s_a2w + 64;     // [128] next block
  float* s_l1bb = s_l1ab + 128;   // [64]
  float* s_S1 = s_l1bb + 64;      // [128][8]
  float* s_S2 = s_S1 + 1024;      // [8][128]
  float* s_part = s_S2 + 1024;    // [4][128]
  float* s_r8 = s_part + 512;     // [4][8][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_r8 + 4096);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int my_row = quad * 32 + lane;
  const int n_rows = a.n_nodes;
  const int n_tiles = (n_rows + TM - 1) / TM;
  const bool has_next = a.nL1a_img != nullptr;

  if (warp == 0) {
    tmem_alloc(s_tmem, TMEM_COLS);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (!a.first) {
    if (tid < 128) {
      s_l2ab[tid] = __ldg(a.w.l2ab + tid);
      s_l2bb[tid] = __ldg(a.w.l2bb + tid);
      s_linb[tid] = __ldg(a.w.linb + tid);
    }
    if (tid < 64) {
      s_a1b[tid] = __ldg(a.w.a1b + tid);
      s_a2w[tid] = __ldg(a.w.a2w + tid);
    }
    for (int i = tid; i < 1024; i += TCN_THREADS) {
      s_S1[i] = __ldg(a.w.S1 + i);
      s_S2[i] = __ldg(a.w.S2 + i);
    }
  }
  if (has_next) {
    if (tid < 128) s_l1ab[tid] = __ldg(a.nl1ab + tid);
    if (tid < 64) s_l1bb[tid] = __ldg(a.nl1bb + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  NodeCtx cx;
  cx.wbuf = base; cx.bars = bars; cx.tmem = *s_tmem;
  cx.trow = cx.tmem + (static_cast<uint32_t>(quad * 32) << 16);
  cx.w_phase = 0; cx.m_phase = 0; cx.tid = tid;
  const float beta_act = a.first ? 1.f : __ldg(a.w.sc + 2);
  const float a2b = a.first ? 0.f : __ldg(a.w.sc + 3);
  constexpr uint32_t IMG_128x128 = 2 * 128 * 128 * 4, IMG_HALF = 2 * 64 * 128 * 4;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = static_cast<int64_t>(tile) * TM + my_row;
    const bool valid = r < n_rows;
    float hnew[32];   // this thread's 32 columns of the updated node state
    if (a.first) {
      // h = embedding[z]
      const int z = valid ? __ldg(a.atom_type + r) : 0;
      const float4* pe = reinterpret_cast<const float4*>(a.emb + (size_t)z * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = __ldg(pe + q);
        hnew[q * 4] = v.x; hnew[q * 4 + 1] = v.y; hnew[q * 4 + 2] = v.z; hnew[q * 4 + 3] = v.w;
      }
      if (tid == 0 && has_next) cx.stream(a.nL1a_img, IMG_128x128);
    } else {
      if (tid == 0) cx.stream(a.w.tL2a, IMG_128x128);
      const bool has_in = valid && __ldg(a.in_ptr + r + 1) > __ldg(a.in_ptr + r);
      // ---- 1. A = agg[:, :128]; conv1.lin2 (+BN) -> t1 = SSP
      {
        const float4* pa = reinterpret_cast<const float4*>(a.agg + (valid ? r : 0) * 192 + part * 32);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float t[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = has_in ? __ldg(pa + c * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            t[q * 4] = v.x; t[q * 4 + 1] = v.y; t[q * 4 + 2] = v.z; t[q * 4 + 3] = v.w;
          }
          st_split16(cx.trow, part * 32 + c * 16, t);
        }
      }
      cx.layer(128, 128);
      if (tid == 0) cx.stream(a.w.tLINa, IMG_128x128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16], t[16];
        const int n0 = part * 32 + c * 16;
        tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = ssp(v[j] + s_l2ab[n0 + j], beta_act);
        st_split16(cx.trow, n0, t);
      }
      // ---- 2. first half of lin: xp = t1 . LIN[0:128]
      cx.layer(128, 128);
      if (tid == 0) cx.stream(a.w.tL2b, IMG_HALF);
      float xp[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16];
        tmem_ld16_acc(cx.trow, part * 32 + c * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) xp[c * 16 + j] = v[j];
      }
      // ---- 3. A[:, :64] = agg[:, 128:192]; conv2.lin2 (+BN) -> t2 = SSP
      if (part < 2) {
        const float4* pa = reinterpret_cast<const float4*>(a.agg + (valid ? r : 0) * 192 + 128 + part * 32);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float t[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = has_in ? __ldg(pa + c * 4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            t[q * 4] = v.x; t[q * 4 + 1] = v.y; t[q * 4 + 2] = v.z; t[q * 4 + 3] = v.w;
          }
          st_split16(cx.trow, part * 32 + c * 16, t);
        }
      }
      cx.layer(64, 128);
      if (tid == 0) cx.stream(a.w.tLINb, IMG_128x128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16], t[16];
        const int n0 = part * 32 + c * 16;
        tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = ssp(v[j] + s_l2bb[n0 + j], beta_act);
        st_split16(cx.trow, n0, t);
      }
      // ---- 4. second half of lin: xc = xp + t2 . LIN[128:256] + b
      cx.layer(128, 128);
      if (tid == 0) cx.stream(a.w.tA1, IMG_HALF);
      float xc[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16], t[16];
        const int n0 = part * 32 + c * 16;
        tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          xc[c * 16 + j] = (xp[c * 16 + j] + v[j]) + s_linb[n0 + j];
          t[j] = xc[c * 16 + j];
        }
        st_split16(cx.trow, n0, t);
      }
      // ---- 5. attention gate
      cx.layer(128, 64);
      if (tid == 0 && has_next) cx.stream(a.nL1a_img, IMG_128x128);
      {
        float v[16];
        const int n0 = part * 16;
        tmem_ld16_acc(cx.trow, n0, v);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fmaf(relu_(v[j] + s_a1b[n0 + j]), s_a2w[n0 + j], acc);
        s_part[part * 128 + my_row] = acc;
      }
      __syncthreads();
      const float gate = sigmoidf_(((s_part[my_row] + s_part[128 + my_row]) + (s_part[256 + my_row] + s_part[384 + my_row])) + a2b);
      // ---- 6. adaptive scaling: r8 = relu(S1^T y), s = sigmoid(S2^T r8), out = y * s
      float r8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r8[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        xc[k] *= gate;   // y
#pragma unroll
        for (int j = 0; j < 8; ++j) r8[j] = fmaf(xc[k], s_S1[(part * 32 + k) * 8 + j], r8[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_r8[(part * 8 + j) * 128 + my_row] = r8[j];
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        r8[j] = relu_((s_r8[(0 * 8 + j) * 128 + my_row] + s_r8[(1 * 8 + j) * 128 + my_row]) +
                      (s_r8[(2 * 8 + j) * 128 + my_row] + s_r8[(3 * 8 + j) * 128 + my_row]));
      const float4* ph = reinterpret_cast<const float4*>(a.h + (valid ? r : 0) * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 ho = valid ? ph[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float hv[4] = {ho.x, ho.y, ho.z, ho.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = q * 4 + u;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) s = fmaf(r8[j], s_S2[j * 128 + part * 32 + k], s);
          hnew[k] = hv[u] + xc[k] * sigmoidf_(s);
        }
      }
    }
    // ---- 7. write h
    if (valid) {
      float4* ph = reinterpret_cast<float4*>(a.h + r * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) ph[q] = make_float4(hnew[q * 4], hnew[q * 4 + 1], hnew[q * 4 + 2], hnew[q * 4 + 3]);
    }
    // ---- 8. next block's x = LeakyReLU(BN(lin1(h)))
    if (has_next) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = hnew[c * 16 + j];
        st_split16(cx.trow, part * 32 + c * 16, t);
      }
      cx.layer(128, 128);
      if (tid == 0) cx.stream(a.nL1b_img, IMG_HALF);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16];
        const int n0 = part * 32 + c * 16;
        tmem_ld16_acc(cx.trow, n0, v);
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(a.xcat + r * 192 + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(leaky02(v[q * 4] + s_l1ab[n0 + q * 4]), leaky02(v[q * 4 + 1] + s_l1ab[n0 + q * 4 + 1]),
                                 leaky02(v[q * 4 + 2] + s_l1ab[n0 + q * 4 + 2]), leaky02(v[q * 4 + 3] + s_l1ab[n0 + q * 4 + 3]));
        }
      }
      cx.layer(128, 64);   // A still holds h
      {
        float v[16];
        const int n0 = part * 16;
        tmem_ld16_acc(cx.trow, n0, v);
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(a.xcat + r * 192 + 128 + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(leaky02(v[q * 4] + s_l1bb[n0 + q * 4]), leaky02(v[q * 4 + 1] + s_l1bb[n0 + q * 4 + 1]),
                                 leaky02(v[q * 4 + 2] + s_l1bb[n0 + q * 4 + 2]), leaky02(v[q * 4 + 3] + s_l1bb[n0 + q * 4 + 3]));
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(cx.tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ GIN layer
// m_i = sum_{e: dst_e = i} relu(x[src_e] + ea_e) + (1 + eps) x_i ;  x' = [relu](BN(Lin(relu(Lin(m))))) + x     gin.py:38-69,112-148
// The gather is row-per-thread (each thread accumulates 32 columns of its node over the CSC segment, two edges in
// flight), so 512 threads keep ~16 K loads outstanding instead of one warp walking a node's edges one by one.
struct TcGinArgs {
  GinW w;
  const float *tG1, *tG2;
  const float* x_in;
  float* x_out;
  const float* ea;
  const int *src, *in_ptr;
  const int* ea_idx;  // row of `ea` for each CSC local edge (pair mode: both directions of a bond share one row), or nullptr
  int n_nodes;
  int last;
  unsigned long long* timing;   // diagnostics (-DAGD_F16_TIMING): [57 + phase] cycles of CTA 0 / thread 0, summed over launches
};

#ifdef AGD_F16_TIMING
#define GIN_TICK(i)                                                                   \
  if (a.timing != nullptr && tid == 0 && blockIdx.x == 0) {                           \
    const long long t_ = clock64();                                                   \
    atomicAdd(a.timing + 57 + (i), static_cast<unsigned long long>(t_ - t_prev));     \
    t_prev = t_;                                                                      \
  }
#else
#define GIN_TICK(i)
#endif

constexpr int GIN_LD = 132;   // padded row stride of the gathered message tile
constexpr size_t TC_GIN_SMEM = 1024 + 131072 + (TM * GIN_LD + 256) * sizeof(float) + 256;

// one lane's float4 of the edge_attr row of local edge e: streamed (evict-first) when every edge has its own row, through the
// pair map with default caching when a row serves both directions of a bond (the second reader may still find it in L2)
template <bool PAIRS>
__device__ __forceinline__ float4 ldg_ea(const float* ea, const int* ea_idx, int e, int lane) {
  if (!PAIRS) return __ldcs(reinterpret_cast<const float4*>(ea + (size_t)e * HID) + lane);
  return __ldg(reinterpret_cast<const float4*>(ea + (size_t)__ldg(ea_idx + e) * HID) + lane);
}

template <bool PAIRS>
__global__ void __launch_bounds__(TCN_THREADS, 1) tc_gin_kernel(const TcGinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS, not generic LD)
  float* s_tile = reinterpret_cast<float*>(base + 131072);   // [128][GIN_LD] gathered messages, row-major
  float* s_g1b = s_tile + TM * GIN_LD;
  float* s_g2b = s_g1b + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_g2b + 128);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef AGD_F16_TIMING
  long long t_prev = clock64();
#endif
  const int quad = warp & 3, part = warp >> 2;
  const int my_row = quad * 32 + lane;
  const int n_rows = a.n_nodes;
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
  if (tid < 128) {
    s_g1b[tid] = __ldg(a.w.g1b + tid);
    s_g2b[tid] = __ldg(a.w.g2b + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  NodeCtx cx;
  cx.wbuf = base; cx.bars = bars; cx.tmem = *s_tmem;
  const bool tmem_at_0 = cx.tmem == 0u;   // always, for the SM's only CTA allocating all 512 columns: the elected-lane issue relies on it
  cx.trow = cx.tmem + (static_cast<uint32_t>(quad * 32) << 16);
  cx.w_phase = 0; cx.m_phase = 0; cx.tid = tid;
  const float ope = __ldg(a.w.sc);
  constexpr uint32_t IMG = 2 * 128 * 128 * 4;
  GIN_TICK(0);   // prologue

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = static_cast<int64_t>(tile) * TM + my_row;
    const bool valid = r < n_rows;
    if (tid == 0) cx.stream(a.tG1, IMG);
    // ---- gather: one warp per node, lanes across the 128 columns (a row = one coalesced 512 B request), 4 edges in flight
    const int64_t row0 = static_cast<int64_t>(tile) * TM;
    for (int rr = warp; rr < TM; rr += TCN_THREADS / 32) {
      const int64_t node = row0 + rr;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (node < n_rows) {
        const float4 self = __ldg(reinterpret_cast<const float4*>(a.x_in + node * HID) + lane);
        const int e0 = __ldg(a.in_ptr + node), e1 = __ldg(a.in_ptr + node + 1);
        int e = e0;
        for (; e + 8 <= e1; e += 8) {   // 8 edges in flight: the gather is latency-bound (L2 / HBM round trips), not bandwidth-bound
          float4 xv[8], ev[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            xv[u] = __ldg(reinterpret_cast<const float4*>(a.x_in + (size_t)__ldg(a.src + e + u) * HID) + lane);
            ev[u] = ldg_ea<PAIRS>(a.ea, a.ea_idx, e + u, lane);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            acc.x += relu_(xv[u].x + ev[u].x); acc.y += relu_(xv[u].y + ev[u].y);
            acc.z += relu_(xv[u].z + ev[u].z); acc.w += relu_(xv[u].w + ev[u].w);
          }
        }
        if (e < e1) {   // remainder (1..7 edges) as one predicated batch instead of one round trip per edge
          const int n = e1 - e;
          float4 xv[8], ev[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < n) {
              xv[u] = __ldg(reinterpret_cast<const float4*>(a.x_in + (size_t)__ldg(a.src + e + u) * HID) + lane);
              ev[u] = ldg_ea<PAIRS>(a.ea, a.ea_idx, e + u, lane);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (u < n) {
              acc.x += relu_(xv[u].x + ev[u].x); acc.y += relu_(xv[u].y + ev[u].y);
              acc.z += relu_(xv[u].z + ev[u].z); acc.w += relu_(xv[u].w + ev[u].w);
            }
          }
        }
        acc.x = fmaf(ope, self.x, acc.x); acc.y = fmaf(ope, self.y, acc.y);
        acc.z = fmaf(ope, self.z, acc.z); acc.w = fmaf(ope, self.w, acc.w);
      }
      *reinterpret_cast<float4*>(s_tile + rr * GIN_LD + lane * 4) = acc;
    }
    GIN_TICK(1);   // own gather
    __syncthreads();
    GIN_TICK(2);   // waiting for the slowest warp
    // ---- each thread lifts its row quarter into TMEM
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float t[16];
      const float4* ps = reinterpret_cast<const float4*>(s_tile + my_row * GIN_LD + part * 32 + c * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = ps[q];
        t[q * 4] = v.x; t[q * 4 + 1] = v.y; t[q * 4 + 2] = v.z; t[q * 4 + 3] = v.w;
      }
      st_split16(cx.trow, part * 32 + c * 16, t);
    }
    GIN_TICK(3);   // lift into TMEM
  
    if (tmem_at_0) cx.layer128_elected(); else cx.layer(128, 128);
    GIN_TICK(4);   // layer 1
    if (tid == 0) cx.stream(a.tG2, IMG);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[16], t[16];
      const int n0 = part * 32 + c * 16;
      tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = relu_(v[j] + s_g1b[n0 + j]);
      st_split16(cx.trow, n0, t);
    }
    GIN_TICK(5);   // epilogue 1
    if (tmem_at_0) cx.layer128_elected(); else cx.layer(128, 128);
    GIN_TICK(6);   // layer 2
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[16];
      const int n0 = part * 32 + c * 16;
      tmem_ld16_acc(cx.trow, n0, v);
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(a.x_out + r * HID + n0);
        const float4* pxo = reinterpret_cast<const float4*>(a.x_in + r * HID + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 xq = __ldg(pxo + q);
          const float xo[4] = {xq.x, xq.y, xq.z, xq.w};
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float t = v[q * 4 + u] + s_g2b[n0 + q * 4 + u];
            if (!a.last) t = relu_(t);
            o[u] = t + xo[u];
          }
          dst[q] = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    GIN_TICK(0);   // epilogue 2 + tile sync (booked with the prologue: the counter block has 7 slots)
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(cx.tmem, TMEM_COLS);
}

void launch_gin_layer_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int layer, const float* x_in, float* x_out) {
  TcGinArgs a{};
  a.w = w.gin[layer];
  a.tG1 = w.gin[layer].tG1; a.tG2 = w.gin[layer].tG2;
  a.x_in = x_in; a.x_out = x_out; a.ea = b.ea_loc; a.src = b.lc_src; a.in_ptr = b.lc_in_ptr; a.ea_idx = b.lc_ea_idx;
  a.n_nodes = b.n_atoms;
  a.timing = c.f16_timing;
  a.last = (layer == c.num_convs_local - 1) ? 1 : 0;
  int tiles = (b.n_atoms + TM - 1) / TM;
  const int grid = tiles < c.num_sms ? tiles : c.num_sms;
  if (a.ea_idx) tc_gin_kernel<true><<<grid, TCN_THREADS, TC_GIN_SMEM, c.stream>>>(a);
  else tc_gin_kernel<false><<<grid, TCN_THREADS, TC_GIN_SMEM, c.stream>>>(a);
  note_launch(c, "gin.layer_tc");
}

void launch_schnet_node_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk) {
  TcNodeArgs a{};
  a.n_nodes = b.n_atoms;
  a.atom_type = b.atom_type;
  a.emb = w.sch_emb;
  a.agg = b.agg;
  a.in_ptr = b.in_ptr;
  a.h = b.h;
  a.xcat = b.xcat;
  a.first = (blk < 0) ? 1 : 0;
  if (blk >= 0) a.w = w.blk[blk];
  const int nxt = blk + 1;
  if (nxt < c.num_convs) {
    a.nL1a_img = w.blk[nxt].tL1a; a.nL1b_img = w.blk[nxt].tL1b;
    a.nl1ab = w.blk[nxt].l1ab; a.nl1bb = w.blk[nxt].l1bb;
  }
  int tiles = (b.n_atoms + TM - 1) / TM;
  const int grid = tiles < c.num_sms ? tiles : c.num_sms;
  tc_node_kernel<<<grid, TCN_THREADS, TC_NODE_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.node_tc");
}

void set_tc_node_attributes() {
  cudaFuncSetAttribute(tc_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_NODE_SMEM);
  cudaFuncSetAttribute(tc_gin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_GIN_SMEM);
  cudaFuncSetAttribute(tc_gin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_GIN_SMEM);
}

}  // namespace agd
