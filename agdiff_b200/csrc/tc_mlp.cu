// MLP edge encoder (edge.py:84-103) and pair MLPs (common.py:86-109) on tcgen05, 3xTF32 with the activation
// operand chained through TENSOR MEMORY (same scheme as tc_filter.cu; algebra as in encoder.cu / pack.py).
//
// One CTA per SM, 512 threads: thread (warp w, lane l) owns tile row 32*(w%4)+l and the column quarter w/4.
// Weights are K-major SWIZZLE_128B [hi|lo] images streamed from L2 by cp.async.bulk into one 128 KB buffer while the
// previous layer's epilogue runs; the pair MLP's 128->64 layer (64 KB) stays resident in a second buffer.
#include "kernels.h"
#include "tc_common.cuh"

namespace agd {

using namespace tc;

constexpr int TCM_THREADS = 512;
constexpr uint32_t IMG128 = 2 * 128 * 128 * 4;   // bytes of a [hi|lo] image of a 128x128 matrix
constexpr uint32_t IMG64 = 2 * 64 * 128 * 4;     // 128 -> 64

// shared scaffolding: TMEM allocation, barriers, the weight stream and the MMA issue loop
struct TcCtx {
  uint8_t* wbuf;
  uint64_t* bars;   // [0] streamed weights landed, [1] mma done, [2] resident weights landed
  uint32_t tmem, trow;
  uint32_t w_phase, m_phase;
  int tid;

  __device__ __forceinline__ void stream(const float* img, uint32_t bytes) {   // tid 0 only
    mbar_expect_tx(&bars[0], bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(img);
    for (uint32_t off = 0; off < bytes; off += 16384) bulk_g2s(wbuf + off, src + off, 16384, &bars[0]);
  }
  // D[128 x N] (+)= A[128 x K] . W^T with W's [hi|lo] image at `b_smem` (hi at +0, lo at +half_bytes); tid 0 only
  // D[128 x N] (+)= A[128 x K] . W^T with W's [hi|lo] image at `b_smem`; tid 0 only.  Cross terms go to the second accumulator.
  __device__ __forceinline__ void issue(uint32_t b_smem, uint32_t /*half_bytes*/, int K, int N, bool accumulate) {
    if (K == 128 && N == 128) issue_3xtf32<128, 128, true>(tmem, b_smem, accumulate);
    else issue_3xtf32<128, 64, true>(tmem, b_smem, accumulate);   // the only other shape used here
    (void)K; (void)N;
    mma_commit(&bars[1]);
  }
  // publish this thread's TMEM stores, run one streamed layer, wait for it (all threads call)
  __device__ __forceinline__ void layer_streamed(int K, int N, bool accumulate) {
    wait_st();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      mbar_wait(&bars[0], w_phase);
      issue(smem_u32(wbuf), static_cast<uint32_t>(K) * N * 4, K, N, accumulate);
    }
    w_phase ^= 1;
  }
  __device__ __forceinline__ void layer_resident(uint32_t b_smem, int K, int N) {
    wait_st();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue(b_smem, static_cast<uint32_t>(K) * N * 4, K, N, false);
    }
  }
  __device__ __forceinline__ void wait_mma() {
    mbar_wait(&bars[1], m_phase);
    m_phase ^= 1;
    fence_after_sync();
  }
};

__device__ __forceinline__ void store_split16(uint32_t addr_hi, uint32_t addr_lo, const float (&t)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split_tf32(t[j], hi[j], lo[j]);
  tmem_st16(addr_hi, hi);
  tmem_st16(addr_lo, lo);
}

// ------------------------------------------------------------------------------------------------ edge encoder
struct TcEncArgs {
  EncW w;
  const float *tW1, *tM2, *tC2;   // [hi|lo] images
  const int* n_rows_dev;
  int n_rows_static;
  const float* e_len;             // global: precomputed lengths
  const int* e_type;
  const float* pos;               // local: lengths from positions
  const int *src, *dst, *canon;
  float *len_csc, *len_canon;
  const float* len_in;            // local: caller-supplied lengths instead of |pos[src]-pos[dst]|
  float* out;                     // g2 (global) / edge_attr (local) [rows][128]
  uint4* g2h;                     // global, AGD_MODE_F16: pre-split fp16 copy of g2 for the filter kernels (layout: g2h_index)
  float lo_scale;                 // 2^S of the lo' part
  int* range_flag;                // set when |g2| leaves the fp16-split range
};

constexpr size_t TC_ENC_SMEM = 1024 + 131072 + 4 * 128 * sizeof(float) + 256;

template <bool LOCAL>
__global__ void __launch_bounds__(TCM_THREADS, 1) tc_encoder_kernel(const TcEncArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS, not generic LD)
  float* s_few = reinterpret_cast<float*>(base + 131072);   // feature_expansion weight
  float* s_feb = s_few + 128;                               // feature_expansion bias
  float* s_c2b = s_feb + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_c2b + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int my_row = quad * 32 + lane;
  const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows_static;
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
    s_few[tid] = __ldg(a.w.fe_w + tid);
    s_feb[tid] = __ldg(a.w.fe_b + tid);
    s_c2b[tid] = __ldg(a.w.c2b + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  TcCtx cx;
  cx.wbuf = base; cx.bars = bars; cx.tmem = *s_tmem;
  cx.trow = cx.tmem + (static_cast<uint32_t>(quad * 32) << 16);
  cx.w_phase = 0; cx.m_phase = 0; cx.tid = tid;

  if (tid == 0 && blockIdx.x < n_tiles) cx.stream(a.tW1, IMG128);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = static_cast<int64_t>(tile) * TM + my_row;
    const bool valid = r < n_rows;
    int type = 0;
    float d = 0.f;
    if (valid) {
      type = __ldg(a.e_type + r);
      if (LOCAL) {
        const int s = __ldg(a.src + r), q = __ldg(a.dst + r);
        const float dx = a.pos[3 * (size_t)s] - a.pos[3 * (size_t)q];
        const float dy = a.pos[3 * (size_t)s + 1] - a.pos[3 * (size_t)q + 1];
        const float dz = a.pos[3 * (size_t)s + 2] - a.pos[3 * (size_t)q + 2];
        d = a.len_in ? __ldg(a.len_in + r) : sqrtf(dx * dx + dy * dy + dz * dz);
        if (part == 0) {
          a.len_csc[r] = d;
          a.len_canon[__ldg(a.canon + r)] = d;
        }
      } else {
        d = __ldg(a.e_len + r);
      }
    }
    // ---- A = gelu(feature_expansion(d))
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float t[16];
      const int k0 = part * 32 + c * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = gelu_fast(fmaf(s_few[k0 + j], d, s_feb[k0 + j]));
      store_split16(cx.trow + COL_AHI + k0, cx.trow + COL_ALO + k0, t);
    }
    cx.layer_streamed(128, 128, false);                 // edge_feature_mlp.0 (x half)
    cx.wait_mma();
    if (tid == 0) cx.stream(a.tM2, IMG128);
    // ---- g1 = gelu(D + T1[type]) -> A
    const float* T1 = a.w.T1 + type * HID;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[16], t[16];
      const int n0 = part * 32 + c * 16;
      float4 tb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) tb[q] = __ldg(reinterpret_cast<const float4*>(T1 + n0) + q);
      tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        t[q * 4 + 0] = gelu_fast(v[q * 4 + 0] + tb[q].x);
        t[q * 4 + 1] = gelu_fast(v[q * 4 + 1] + tb[q].y);
        t[q * 4 + 2] = gelu_fast(v[q * 4 + 2] + tb[q].z);
        t[q * 4 + 3] = gelu_fast(v[q * 4 + 3] + tb[q].w);
      }
      store_split16(cx.trow + COL_AHI + n0, cx.trow + COL_ALO + n0, t);
    }
    cx.layer_streamed(128, 128, false);                 // combination_mlp.0 o edge_feature_mlp.2
    cx.wait_mma();
    const bool more = tile + static_cast<int>(gridDim.x) < n_tiles;
    if (tid == 0) {
      if (LOCAL) cx.stream(a.tC2, IMG128);
      else if (more) cx.stream(a.tW1, IMG128);
    }
    // ---- g2 = gelu(D + T2[type])
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    const float* T2 = a.w.T2 + type * HID;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[16], t[16];
      const int n0 = part * 32 + c * 16;
      float4 tb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) tb[q] = __ldg(reinterpret_cast<const float4*>(T2 + n0) + q);
      tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        t[q * 4 + 0] = gelu_fast(v[q * 4 + 0] + tb[q].x);
        t[q * 4 + 1] = gelu_fast(v[q * 4 + 1] + tb[q].y);
        t[q * 4 + 2] = gelu_fast(v[q * 4 + 2] + tb[q].z);
        t[q * 4 + 3] = gelu_fast(v[q * 4 + 3] + tb[q].w);
      }
      if (LOCAL) {
        store_split16(cx.trow + COL_AHI + n0, cx.trow + COL_ALO + n0, t);
      } else if (valid) {
        float4* dst = reinterpret_cast<float4*>(a.out + r * HID + n0);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(t[q * 4], t[q * 4 + 1], t[q * 4 + 2], t[q * 4 + 3]);
        if (a.g2h) {   // the same 16 features as packed fp16 hi / lo' pairs: words n0/2 .. n0/2+7 of the row and of its lo' half
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split2_f16(t[2 * j], t[2 * j + 1], a.lo_scale, hi[j], lo[j], amax);
          const int w4 = n0 >> 3;
          a.g2h[g2h_index(r, w4)] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          a.g2h[g2h_index(r, w4 + 1)] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          a.g2h[g2h_index(r, 16 + w4)] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          a.g2h[g2h_index(r, 16 + w4 + 1)] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
    if (!LOCAL && f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
    if (LOCAL) {
      cx.layer_streamed(128, 128, false);               // combination_mlp.2
      cx.wait_mma();
      if (tid == 0 && more) cx.stream(a.tW1, IMG128);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[16];
        const int n0 = part * 32 + c * 16;
        tmem_ld16_acc(cx.trow, n0, v);
        if (valid) {
          float4* dst = reinterpret_cast<float4*>(a.out + r * HID + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(v[q * 4] + s_c2b[n0 + q * 4], v[q * 4 + 1] + s_c2b[n0 + q * 4 + 1],
                                 v[q * 4 + 2] + s_c2b[n0 + q * 4 + 2], v[q * 4 + 3] + s_c2b[n0 + q * 4 + 3]);
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

// ------------------------------------------------------------------------------------------------ pair MLP
struct TcPairArgs {
  PairW w;
  const float *tP1h, *tP1e, *tP2;   // [hi|lo] images: 128x128, 128x128, 64x128
  const int* n_rows_dev;
  int n_rows_static;
  const float* h;      // node features [N][128]
  const float* feat;   // g2 (global) / edge_attr (local) [rows][128]
  const int *src, *dst, *canon;
  float *s_csc, *s_canon;
};

constexpr size_t TC_PAIR_SMEM = 1024 + 131072 + 65536 + (128 + 64 + 64 + 4 * 128) * sizeof(float) + 256;

__global__ void __launch_bounds__(TCM_THREADS, 1) tc_pair_kernel(const TcPairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer (LDS, not generic LD)
  uint8_t* w2buf = base + 131072;                              // resident 128->64 layer, 64 KB
  float* s_p1b = reinterpret_cast<float*>(base + 131072 + 65536);
  float* s_p2b = s_p1b + 128;
  float* s_p3w = s_p2b + 64;
  float* s_part = s_p3w + 64;                                  // [4][128] partial scores per column quarter
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 4 * 128);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int my_row = quad * 32 + lane;
  const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows_static;
  const int n_tiles = (n_rows + TM - 1) / TM;

  if (warp == 0) {
    tmem_alloc(s_tmem, TMEM_COLS);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (tid < 128) s_p1b[tid] = __ldg(a.w.p1b + tid);
  if (tid < 64) {
    s_p2b[tid] = __ldg(a.w.p2b + tid);
    s_p3w[tid] = __ldg(a.w.p3w + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  TcCtx cx;
  cx.wbuf = base; cx.bars = bars; cx.tmem = *s_tmem;
  cx.trow = cx.tmem + (static_cast<uint32_t>(quad * 32) << 16);
  cx.w_phase = 0; cx.m_phase = 0; cx.tid = tid;
  const float p3b = __ldg(a.w.p3b);

  if (tid == 0 && blockIdx.x < n_tiles) {
    mbar_expect_tx(&bars[2], IMG64);
    for (uint32_t off = 0; off < IMG64; off += 16384)
      bulk_g2s(w2buf + off, reinterpret_cast<const uint8_t*>(a.tP2) + off, 16384, &bars[2]);
    cx.stream(a.tP1h, IMG128);
    mbar_wait(&bars[2], 0);     // resident layer landed before its first use (tid 0 is the only MMA issuer)
  }

  // products h[src]*h[dst] of this thread's 32 columns, one tile ahead
  float4 hh[8];
  auto prefetch_hh = [&](int t) {
    const int64_t rr = static_cast<int64_t>(t) * TM + my_row;
    if (t < n_tiles && rr < n_rows) {
      const float4* ps = reinterpret_cast<const float4*>(a.h + (size_t)__ldg(a.src + rr) * HID + part * 32);
      const float4* pd = reinterpret_cast<const float4*>(a.h + (size_t)__ldg(a.dst + rr) * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = __ldg(ps + q), y = __ldg(pd + q);
        hh[q] = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) hh[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  prefetch_hh(blockIdx.x);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r = static_cast<int64_t>(tile) * TM + my_row;
    const bool valid = r < n_rows;
    // ---- A = h[row] * h[col]
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float t[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        t[q * 4 + 0] = hh[c * 4 + q].x; t[q * 4 + 1] = hh[c * 4 + q].y; t[q * 4 + 2] = hh[c * 4 + q].z; t[q * 4 + 3] = hh[c * 4 + q].w;
      }
      store_split16(cx.trow + COL_AHI + part * 32 + c * 16, cx.trow + COL_ALO + part * 32 + c * 16, t);
    }
    cx.layer_streamed(128, 128, false);                 // layers.0, h half
    // the edge-feature rows of this tile travel while the tensor core works
    float4 fe[8];
    if (valid) {
      const float4* pf = reinterpret_cast<const float4*>(a.feat + r * HID + part * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) fe[q] = __ldg(pf + q);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) fe[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cx.wait_mma();
    if (tid == 0) cx.stream(a.tP1e, IMG128);
    // ---- A = edge features (the first product is complete, A may be overwritten; D keeps accumulating)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float t[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        t[q * 4 + 0] = fe[c * 4 + q].x; t[q * 4 + 1] = fe[c * 4 + q].y; t[q * 4 + 2] = fe[c * 4 + q].z; t[q * 4 + 3] = fe[c * 4 + q].w;
      }
      store_split16(cx.trow + COL_AHI + part * 32 + c * 16, cx.trow + COL_ALO + part * 32 + c * 16, t);
    }
    cx.layer_streamed(128, 128, true);                  // layers.0, edge half, accumulated
    prefetch_hh(tile + static_cast<int>(gridDim.x));
    cx.wait_mma();
    if (tid == 0 && tile + static_cast<int>(gridDim.x) < n_tiles) cx.stream(a.tP1h, IMG128);
    // ---- r1 = relu(D + b1) -> A
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float v[16], t[16];
      const int n0 = part * 32 + c * 16;
      tmem_ld16_acc(cx.trow, n0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = relu_(v[j] + s_p1b[n0 + j]);
      store_split16(cx.trow + COL_AHI + n0, cx.trow + COL_ALO + n0, t);
    }
    cx.layer_resident(smem_u32(w2buf), 128, 64);        // layers.1
    cx.wait_mma();
    // ---- score = layers.2(relu(D + b2)): 16 columns per thread, four quarters combined through smem
    {
      float v[16];
      const int n0 = part * 16;
      tmem_ld16_acc(cx.trow, n0, v);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = fmaf(relu_(v[j] + s_p2b[n0 + j]), s_p3w[n0 + j], acc);
      s_part[part * 128 + my_row] = acc;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (part == 0 && valid) {
      const float s = ((s_part[my_row] + s_part[128 + my_row]) + (s_part[256 + my_row] + s_part[384 + my_row])) + p3b;
      a.s_csc[r] = s;
      a.s_canon[__ldg(a.canon + r)] = s;
    }
    // s_part is rewritten only after the next tile's barriers
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(cx.tmem, TMEM_COLS);
}

static int tc_grid(int64_t rows_cap, int num_sms) {
  int64_t t = (rows_cap + TM - 1) / TM;
  if (t < 1) t = 1;
  return (int)(t < num_sms ? t : num_sms);
}

void launch_encoder_global_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  TcEncArgs a{};
  a.w = w.enc; a.tW1 = w.tenc_W1; a.tM2 = w.tenc_M2; a.tC2 = w.tenc_C2;
  a.n_rows_dev = b.counters;
  a.e_len = b.e_len; a.e_type = b.e_type; a.out = b.g2;
  if (c.use_tc == 2) {
    a.g2h = b.g2h;
    a.lo_scale = static_cast<float>(1 << f16_lo_shift());
    a.range_flag = b.counters + 4;
  }
  tc_encoder_kernel<false><<<tc_grid(b.cap, c.num_sms), TCM_THREADS, TC_ENC_SMEM, c.stream>>>(a);
  note_launch(c, "encoder.global_tc");
}

void launch_encoder_local_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos) {
  if (b.n_local == 0) return;
  TcEncArgs a{};
  a.w = w.enc; a.tW1 = w.tenc_W1; a.tM2 = w.tenc_M2; a.tC2 = w.tenc_C2;
  a.n_rows_dev = nullptr; a.n_rows_static = b.n_local;
  a.e_type = b.lc_type; a.pos = pos; a.src = b.lc_src; a.dst = b.lc_dst; a.canon = b.lc_canon;
  a.len_csc = b.lc_len; a.len_canon = b.lcc_len; a.len_in = b.lc_len_in; a.out = b.ea_loc;
  tc_encoder_kernel<true><<<tc_grid(b.n_local, c.num_sms), TCM_THREADS, TC_ENC_SMEM, c.stream>>>(a);
  note_launch(c, "encoder.local_tc");
}

void launch_pair_global_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  TcPairArgs a{};
  a.w = w.pg; a.tP1h = w.tpg_P1h; a.tP1e = w.tpg_P1e; a.tP2 = w.tpg_P2;
  a.n_rows_dev = b.counters;
  a.h = b.h; a.feat = b.g2; a.src = b.e_src; a.dst = b.e_dst; a.canon = b.e_canon;
  a.s_csc = b.s_csc; a.s_canon = b.s_canon;
  tc_pair_kernel<<<tc_grid(b.cap, c.num_sms), TCM_THREADS, TC_PAIR_SMEM, c.stream>>>(a);
  note_launch(c, "pair.global_tc");
}

void launch_pair_local_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* h_local) {
  if (b.n_local == 0) return;
  TcPairArgs a{};
  a.w = w.pl; a.tP1h = w.tpl_P1h; a.tP1e = w.tpl_P1e; a.tP2 = w.tpl_P2;
  a.n_rows_dev = nullptr; a.n_rows_static = b.n_local;
  a.h = h_local; a.feat = b.ea_loc; a.src = b.lc_src; a.dst = b.lc_dst; a.canon = b.lc_canon;
  a.s_csc = b.sl_csc; a.s_canon = b.sl_canon;
  tc_pair_kernel<<<tc_grid(b.n_local, c.num_sms), TCM_THREADS, TC_PAIR_SMEM, c.stream>>>(a);
  note_launch(c, "pair.local_tc");
}

void set_tc_mlp_attributes() {
  cudaFuncSetAttribute(tc_encoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_ENC_SMEM);
  cudaFuncSetAttribute(tc_encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_ENC_SMEM);
  cudaFuncSetAttribute(tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_PAIR_SMEM);
}

}  // namespace agd
