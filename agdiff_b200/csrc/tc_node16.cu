// Node-side chain of a SchNet interaction block on tcgen05 with fp16-split operands (numerics: tc16_common.cuh / tc_filter16.cu;
// algebra and thread mapping as in tc_node.cu):
//
//   t1 = SSP(BN(lin2_1(agg[:, :128])))            t2 = SSP(BN(lin2_2(agg[:, 128:])))           schnet.py:157-158,206-207
//   xc = lin([t1, t2])  (K = 256 as two K = 128 passes; the first partial waits in registers)     schnet.py:208
//   gate = sigmoid(a2 . relu(A1 xc + a1b) + a2b);  y = xc * gate                                 schnet.py:211-214
//   s = sigmoid(S2^T relu(S1^T y));  h += y * s                                                  schnet.py:230-234,278-280
//   next block: x = LeakyReLU_0.2(BN(lin1(h))) for conv1 (128) and conv2 (64)                    schnet.py:152-155
//
// One CTA per SM, 512 threads: thread (warp w, lane l) owns node row 32*(w%4)+l and the column quarter w/4.  The chain has up to
// seven layers with seven different weight matrices (352 KB as fp16 hi/lo' images), so they are streamed - but, at half the
// TF32 size, through TWO 64 KB buffers: the image of layer i+2 is requested the moment layer i has completed and travels during
// the whole of layer i+1, instead of being waited for inside every layer as in the 3xTF32 kernel.  kind::f16 halves the MMA time.
#include "kernels.h"
#include "tc16_common.cuh"

namespace agd {

using namespace tc;

constexpr int TCN16_THREADS = 512;
constexpr uint32_t NIMG_128 = 2u * 128u * 128u * 2u;   // bytes of a [hi | lo'] fp16 image, 128 x 128
constexpr uint32_t NIMG_HALF = 2u * 64u * 128u * 2u;   // 128 -> 64 or K = 64 -> 128

struct TcNode16Args {
  BlkW w;                                   // block whose convs just aggregated (unused when first)
  const float *hL2a, *hLINa, *hL2b, *hLINb, *hA1, *nsc;   // this block's fp16 images + inverse scales [L2a, LINa, L2b, LINb, A1]
  const float *nhL1a, *nhL1b, *nnsc;        // NEXT block's lin1 images + inverse scales [L1a, L1b] (nullptr after the last block)
  const float *nl1ab, *nl1bb;
  const float* emb;
  const int* atom_type;
  int n_nodes;
  int first;
  const float* agg;   // [N][192]
  const int* in_ptr;  // [N+1]: atoms without in-edges aggregate to zero (their agg rows are not written by the fused kernels)
  float* h;           // [N][128]
  float* xcat;        // [N][192]
  int scaled;
  int* range_flag;
  unsigned long long* timing;   // diagnostics (-DAGD_F16_TIMING): [40 + phase] cycles of CTA 0 / thread 0, summed over launches
};

#ifdef AGD_F16_TIMING
#define NODE_TICK(i)                                                                  \
  if (a.timing != nullptr && tid == 0 && blockIdx.x == 0) {                           \
    const long long t_ = clock64();                                                   \
    atomicAdd(a.timing + 40 + (i), static_cast<unsigned long long>(t_ - t_prev));     \
    t_prev = t_;                                                                      \
  }
#else
#define NODE_TICK(i)
#endif

constexpr size_t TC_NODE16_SMEM = 1024 + 2 * NIMG_128 + (128 * 3 + 64 * 2 + 128 + 64 + 1024 + 1024 + 512 + 4096) * sizeof(float) + 256;

struct Node16Ctx {
  // (no arrays indexed by a run-time value in here: they would put the whole struct into local memory - with ~70 KB of L1 left
  // that is an L2 round trip for the phase bit and the buffer address in front of every layer's MMA issue)
  uint8_t *wbuf0, *wbuf1;
  uint64_t* bars;      // [0], [1]: weights landed in buffer 0 / 1, [2]: mma done
  uint32_t tmem, trow;
  uint32_t wph0, wph1, m_phase;
  int tid;
  int n_streamed, n_layers;   // running counters: layer i uses buffer i & 1
  bool scaled;

  // request the image of the NEXT not-yet-requested layer (tid 0 only); the buffer's previous user completed two layers ago
  __device__ __forceinline__ void stream(const float* img, uint32_t bytes) {
    const int b = n_streamed & 1;
    uint8_t* dst = b ? wbuf1 : wbuf0;
    mbar_expect_tx(&bars[b], bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img);
    for (uint32_t off = 0; off < bytes; off += 16384) bulk_g2s(dst + off, src + off, 16384, &bars[b]);
    ++n_streamed;
  }
  template <int K, int N>
  __device__ __forceinline__ void layer() {   // all threads
    wait_st();
    fence_before_sync();
    __syncthreads();
    const int b = n_layers & 1;
    if (tid < 32) {   // warp 0 (converged): waits for the layer's weights, one elected lane issues the layer
      fence_after_sync();
      mbar_wait(&bars[b], b ? wph1 : wph0);
      if (elect_one()) {
        issue_3xf16_ct<K, N, 0u>(smem_u32(b ? wbuf1 : wbuf0), static_cast<uint32_t>(K) * N * 2u, scaled);
        mma_commit(&bars[2]);
      }
      __syncwarp();
    }
    if (b) wph1 ^= 1u; else wph0 ^= 1u;
    ++n_layers;
    mbar_wait(&bars[2], m_phase);
    m_phase ^= 1u;
    fence_after_sync();
  }
};

// 32 fp32 values of one row (this thread's column quarter) -> 16 hi words + 16 lo' words at operand columns [col, col + 16)
__device__ __forceinline__ void node_store_split32(uint32_t trow, int col, const float (&t)[32], float lo_scale, __half2& amax) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split2_f16(t[2 * j], t[2 * j + 1], lo_scale, hi[j], lo[j], amax);
  tmem_st16(trow + C16_AHI + col, hi);
  tmem_st16(trow + C16_ALO + col, lo);
}

__global__ void __launch_bounds__(TCN16_THREADS, 1) tc_node16_kernel(const TcNode16Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* s_l2ab = reinterpret_cast<float*>(base + 2 * NIMG_128);
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
#ifdef AGD_F16_TIMING
  long long t_prev = clock64();
#endif
  const int quad = warp & 3, part = warp >> 2;
  const int my_row = quad * 32 + lane;
  const int n_rows = a.n_nodes;
  const int n_tiles = (n_rows + TM - 1) / TM;
  const bool has_next = a.nhL1a != nullptr;

  if (warp == 0) {
    tmem_alloc(s_tmem, 256);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
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
    for (int i = tid; i < 1024; i += TCN16_THREADS) {
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
  Node16Ctx cx;
  cx.wbuf0 = base; cx.wbuf1 = base + NIMG_128; cx.bars = bars; cx.tmem = *s_tmem;
  if (cx.tmem != 0u) {   // the MMA issue uses compile-time TMEM addresses: this CTA is alone on its SM, so the allocation starts at 0
    if (tid == 0) atomicOr(a.range_flag, 2);   // (if it ever does not, the host re-runs the call in mode 1)
    __syncthreads();
    if (warp == 0) tmem_dealloc(cx.tmem, 256);
    return;
  }
  cx.trow = cx.tmem + (static_cast<uint32_t>(quad * 32) << 16);
  cx.wph0 = cx.wph1 = 0; cx.m_phase = 0; cx.tid = tid; cx.n_streamed = 0; cx.n_layers = 0;
  cx.scaled = a.scaled != 0;
  const float beta_act = a.first ? 1.f : __ldg(a.w.sc + 2);
  const float a2b = a.first ? 0.f : __ldg(a.w.sc + 3);
  const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
  float iL2a = 1.f, iLINa = 1.f, iL2b = 1.f, iLINb = 1.f, iA1 = 1.f, iL1a = 1.f, iL1b = 1.f;
  if (!a.first) {
    iL2a = __ldg(a.nsc + 0); iLINa = __ldg(a.nsc + 1); iL2b = __ldg(a.nsc + 2); iLINb = __ldg(a.nsc + 3); iA1 = __ldg(a.nsc + 4);
  }
  if (has_next) {
    iL1a = __ldg(a.nnsc + 0); iL1b = __ldg(a.nnsc + 1);
  }
  __half2 amax = __floats2half2_rn(0.f, 0.f);
  NODE_TICK(0);   // prologue

  // the first two layers of the first tile
  if (tid == 0 && static_cast<int>(blockIdx.x) < n_tiles) {
    if (a.first) {
      if (has_next) {
        cx.stream(a.nhL1a, NIMG_128);
        cx.stream(a.nhL1b, NIMG_HALF);
      }
    } else {
      cx.stream(a.hL2a, NIMG_128);
      cx.stream(a.hLINa, NIMG_128);
    }
  }

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = static_cast<int64_t>(tile) * TM + my_row;
    const bool valid = r < n_rows;
    const bool more = tile + static_cast<int>(gridDim.x) < n_tiles;
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
    } else {
      const bool has_in = valid && __ldg(a.in_ptr + r + 1) > __ldg(a.in_ptr + r);
      // ---- 1. A = agg[:, :128]; conv1.lin2 (+BN) -> t1 = SSP
      {
        const float4* pa = reinterpret_cast<const float4*>(a.agg + (valid ? r : 0) * 192 + part * 32);
        float t[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = has_in ? __ldg(pa + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[q * 4] = v.x; t[q * 4 + 1] = v.y; t[q * 4 + 2] = v.z; t[q * 4 + 3] = v.w;
        }
        node_store_split32(cx.trow, part * 16, t, lo_scale, amax);
      }
      NODE_TICK(1);   // agg load + split
      cx.layer<128, 128>();                                    // L2a
      NODE_TICK(2);
      if (tid == 0) cx.stream(a.hL2b, NIMG_HALF);
      {
        uint32_t v[32];
        float t[32];
        const int n0 = part * 32;
        tmem_ld32(cx.trow + C16_D + n0, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = ssp_fast(fmaf(__uint_as_float(v[j]), iL2a, s_l2ab[n0 + j]), beta_act);
        node_store_split32(cx.trow, part * 16, t, lo_scale, amax);
      }
      // ---- 2. first half of lin: xp = t1 . LIN[0:128]
      NODE_TICK(3);   // epilogue L2a
      cx.layer<128, 128>();                                    // LINa
      NODE_TICK(4);
      if (tid == 0) cx.stream(a.hLINb, NIMG_128);
      float xp[32];
      {
        uint32_t v[32];
        tmem_ld32(cx.trow + C16_D + part * 32, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) xp[j] = __uint_as_float(v[j]) * iLINa;
      }
      // ---- 3. A[:, :64] = agg[:, 128:192]; conv2.lin2 (+BN) -> t2 = SSP
      if (part < 2) {
        const float4* pa = reinterpret_cast<const float4*>(a.agg + (valid ? r : 0) * 192 + 128 + part * 32);
        float t[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = has_in ? __ldg(pa + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[q * 4] = v.x; t[q * 4 + 1] = v.y; t[q * 4 + 2] = v.z; t[q * 4 + 3] = v.w;
        }
        node_store_split32(cx.trow, part * 16, t, lo_scale, amax);
      }
      NODE_TICK(5);   // xp + agg2 load
      cx.layer<64, 128>();                                     // L2b (K = 64)
      NODE_TICK(6);
      if (tid == 0) cx.stream(a.hA1, NIMG_HALF);
      {
        uint32_t v[32];
        float t[32];
        const int n0 = part * 32;
        tmem_ld32(cx.trow + C16_D + n0, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = ssp_fast(fmaf(__uint_as_float(v[j]), iL2b, s_l2bb[n0 + j]), beta_act);
        node_store_split32(cx.trow, part * 16, t, lo_scale, amax);
      }
      // ---- 4. second half of lin: xc = xp + t2 . LIN[128:256] + b
      NODE_TICK(7);   // epilogue L2b
      cx.layer<128, 128>();                                    // LINb
      NODE_TICK(8);
      if (tid == 0) {
        if (has_next) cx.stream(a.nhL1a, NIMG_128);
        else if (more) cx.stream(a.hL2a, NIMG_128);
      }
      float xc[32];
      {
        uint32_t v[32];
        const int n0 = part * 32;
        tmem_ld32(cx.trow + C16_D + n0, v);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) xc[j] = (xp[j] + __uint_as_float(v[j]) * iLINb) + s_linb[n0 + j];
        node_store_split32(cx.trow, part * 16, xc, lo_scale, amax);
      }
      // ---- 5. attention gate
      NODE_TICK(9);   // epilogue LINb
      cx.layer<128, 64>();                                     // A1
      NODE_TICK(10);
      if (tid == 0) {
        if (has_next) cx.stream(a.nhL1b, NIMG_HALF);
        else if (more) cx.stream(a.hLINa, NIMG_128);
      }
      // the residual input h: requested here, two block-wide barriers ahead of its use (a load cannot be hoisted across them
      // by the compiler, and with the L1 carved down to almost nothing it is an L2 round trip)
      float4 hold[8];
      {
        const float4* ph = reinterpret_cast<const float4*>(a.h + (valid ? r : 0) * HID + part * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) hold[q] = valid ? ph[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      {
        uint32_t v[16];
        const int n0 = part * 16;
        tmem_ld16(cx.trow + C16_D + n0, v);
        wait_ld();
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fmaf(relu_(fmaf(__uint_as_float(v[j]), iA1, s_a1b[n0 + j])), s_a2w[n0 + j], acc);
        s_part[part * 128 + my_row] = acc;
      }
      __syncthreads();
      const float gate = sigmoid_fast(((s_part[my_row] + s_part[128 + my_row]) + (s_part[256 + my_row] + s_part[384 + my_row])) + a2b);
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
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float hv[4] = {hold[q].x, hold[q].y, hold[q].z, hold[q].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = q * 4 + u;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) s = fmaf(r8[j], s_S2[j * 128 + part * 32 + k], s);
          hnew[k] = hv[u] + xc[k] * sigmoid_fast(s);
        }
      }
    }
    NODE_TICK(11);   // gate + adaptive scaling
    // ---- 7. write h
    if (valid) {
      float4* ph = reinterpret_cast<float4*>(a.h + r * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) ph[q] = make_float4(hnew[q * 4], hnew[q * 4 + 1], hnew[q * 4 + 2], hnew[q * 4 + 3]);
    }
    // ---- 8. next block's x = LeakyReLU(BN(lin1(h)))
    if (has_next) {
      node_store_split32(cx.trow, part * 16, hnew, lo_scale, amax);
      NODE_TICK(12);   // h store + split
      cx.layer<128, 128>();                                    // next L1a
      NODE_TICK(13);
      if (tid == 0 && more) cx.stream(a.first ? a.nhL1a : a.hL2a, NIMG_128);
      {
        uint32_t v[32];
        const int n0 = part * 32;
        tmem_ld32(cx.trow + C16_D + n0, v);
        wait_ld();
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(a.xcat + r * 192 + n0);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(leaky02(fmaf(__uint_as_float(v[q * 4]), iL1a, s_l1ab[n0 + q * 4])),
                                 leaky02(fmaf(__uint_as_float(v[q * 4 + 1]), iL1a, s_l1ab[n0 + q * 4 + 1])),
                                 leaky02(fmaf(__uint_as_float(v[q * 4 + 2]), iL1a, s_l1ab[n0 + q * 4 + 2])),
                                 leaky02(fmaf(__uint_as_float(v[q * 4 + 3]), iL1a, s_l1ab[n0 + q * 4 + 3])));
        }
      }
      NODE_TICK(14);   // epilogue L1a
      cx.layer<128, 64>();   // next L1b; A still holds h
      NODE_TICK(15);
      if (tid == 0 && more) cx.stream(a.first ? a.nhL1b : a.hLINa, a.first ? NIMG_HALF : NIMG_128);
      {
        uint32_t v[16];
        const int n0 = part * 16;
        tmem_ld16(cx.trow + C16_D + n0, v);
        wait_ld();
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(a.xcat + r * 192 + 128 + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(leaky02(fmaf(__uint_as_float(v[q * 4]), iL1b, s_l1bb[n0 + q * 4])),
                                 leaky02(fmaf(__uint_as_float(v[q * 4 + 1]), iL1b, s_l1bb[n0 + q * 4 + 1])),
                                 leaky02(fmaf(__uint_as_float(v[q * 4 + 2]), iL1b, s_l1bb[n0 + q * 4 + 2])),
                                 leaky02(fmaf(__uint_as_float(v[q * 4 + 3]), iL1b, s_l1bb[n0 + q * 4 + 3])));
        }
      }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    NODE_TICK(16);   // epilogue L1b
  }
  if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
  __syncthreads();
  if (warp == 0) tmem_dealloc(cx.tmem, 256);
}

void launch_schnet_node_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk) {
  TcNode16Args a{};
  a.n_nodes = b.n_atoms;
  a.atom_type = b.atom_type;
  a.emb = w.sch_emb;
  a.agg = b.agg;
  a.in_ptr = b.in_ptr;
  a.h = b.h;
  a.xcat = b.xcat;
  a.first = (blk < 0) ? 1 : 0;
  if (blk >= 0) {
    a.w = w.blk[blk];
    a.hL2a = w.blk[blk].hL2a; a.hLINa = w.blk[blk].hLINa; a.hL2b = w.blk[blk].hL2b; a.hLINb = w.blk[blk].hLINb;
    a.hA1 = w.blk[blk].hA1; a.nsc = w.blk[blk].hnsc;
  }
  const int nxt = blk + 1;
  if (nxt < c.num_convs) {
    a.nhL1a = w.blk[nxt].hL1a; a.nhL1b = w.blk[nxt].hL1b; a.nnsc = w.blk[nxt].hnsc + 5;
    a.nl1ab = w.blk[nxt].l1ab; a.nl1bb = w.blk[nxt].l1bb;
  }
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  a.timing = c.f16_timing;
  int tiles = (b.n_atoms + TM - 1) / TM;
  const int grid = tiles < c.num_sms ? tiles : c.num_sms;
  tc_node16_kernel<<<grid, TCN16_THREADS, TC_NODE16_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.node_f16");
}

void set_tc_node16_attributes() {
  cudaFuncSetAttribute(tc_node16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_NODE16_SMEM);
}

}  // namespace agd
