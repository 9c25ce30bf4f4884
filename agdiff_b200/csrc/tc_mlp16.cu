// MLP edge encoder (edge.py:84-103) and pair MLPs (common.py:86-109) on tcgen05 with fp16-split operands and TWO edge tiles
// in flight per SM - the scheme of tc_filter16.cu (numerics, slot geometry, MMA issue: tc16_common.cuh) applied to the other
// per-edge chains.  Algebra as in encoder.cu / pack.py:
//
//   encoder:  a0 = gelu(fe_w d + fe_b);  g1 = gelu(W1 a0 + T1[type]);  g2 = gelu(M2 g1 + T2[type]);  [local: ea = C2 g2 + c2b]
//   pair:     r1 = relu(P1h (h_src * h_dst) + P1e feat + b1);  r2 = relu(P2 r1 + b2);  score = p3 . r2 + b3
//
// One CTA per SM, 512 threads = 2 groups x 8 warps; thread (quad q, lane l, half h) of a group owns tile row 32q+l and one half
// of the columns.  All weight images of a kernel stay resident in shared memory (bulk-copied once per CTA).
#include <cstdlib>

#include "kernels.h"
#include "tc16_common.cuh"

namespace agd {

using namespace tc;

constexpr uint32_t IMG16_128 = 2u * 128u * 128u * 2u;   // bytes of a [hi | lo'] fp16 image of a 128x128 matrix
constexpr uint32_t IMG16_64 = 2u * 64u * 128u * 2u;     // 128 -> 64

// pipeline bookkeeping of one thread of a slot group
struct Slot16 {
  uint32_t slot, trow;
  uint64_t *a_ready, *d_ready;
  uint32_t dph, aph;
  bool issuer, scaled;   // issuer: this thread's WARP issues the group's MMAs (one elected lane)

  // this thread's operand stores are done: publish them, (issuer) run one layer on the tensor core
  template <int K, int N>
  __device__ __forceinline__ void run_layer(uint32_t w_smem, uint32_t half_bytes) {
    wait_st();
    fence_before_sync();
    mbar_arrive(a_ready);
    if (issuer) {   // warp-uniform: the group's first warp
      mbar_wait(a_ready, aph);
      fence_after_sync();
      if (elect_one()) {
        if (slot == 0u) issue_3xf16_ct<K, N, 0u>(w_smem, half_bytes, scaled);
        else issue_3xf16_ct<K, N, SLOT_COLS>(w_smem, half_bytes, scaled);
        mma_commit(d_ready);
      }
      __syncwarp();
    }
    aph ^= 1u;
  }
  __device__ __forceinline__ void wait_layer() {
    mbar_wait(d_ready, dph);
    dph ^= 1u;
    fence_after_sync();
  }
};

// 32 fp32 values of one row -> 16 hi words + 16 lo' words at operand columns [col, col + 16)
__device__ __forceinline__ void store_split32(uint32_t trow, int col, const float (&t)[32], float lo_scale, __half2& amax) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split2_f16(t[2 * j], t[2 * j + 1], lo_scale, hi[j], lo[j], amax);
  tmem_st16(trow + C16_AHI + col, hi);
  tmem_st16(trow + C16_ALO + col, lo);
}

// ------------------------------------------------------------------------------------------------ edge encoder
struct TcEnc16Args {
  EncW w;                            // fe_w, fe_b, T1, T2, c2b
  const uint32_t *hW1, *hM2, *hC2;   // fp16 [hi | lo'] images, 128x128 each
  const float* wsc;                  // inverse power-of-two scales of W1, M2, C2
  const int* n_rows_dev;
  int n_rows_static;
  const float* e_len;                // global: precomputed lengths
  const int* e_type;
  const float* pos;                  // local: lengths from positions
  const int *src, *dst, *canon;
  float *len_csc, *len_canon;
  const float* len_in;               // local: caller-supplied lengths instead of |pos[src]-pos[dst]|
  float* out;                        // local: edge_attr [rows][128]
  uint4* g2h;                        // global: pre-split g2 (tc_common.cuh: g2h_index), what the fp16 consumers read
  float* g2;                         // global: fp32 g2 [rows][128] for a 3xTF32 pair kernel, or nullptr
  int scaled;
  int* range_flag;
};

template <bool LOCAL>
struct TcEnc16Smem {
  static constexpr size_t bytes = 1024 + (LOCAL ? 3 : 2) * IMG16_128 + 3 * 128 * sizeof(float) + 8 * sizeof(uint64_t) + 64;
};

template <bool LOCAL>
__global__ void __launch_bounds__(F16_THREADS, 1) tc_encoder16_kernel(const TcEnc16Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* w1 = base;
  uint8_t* m2 = base + IMG16_128;
  uint8_t* c2 = base + 2 * IMG16_128;
  float* s_few = reinterpret_cast<float*>(base + (LOCAL ? 3 : 2) * IMG16_128);
  float* s_feb = s_few + 128;
  float* s_c2b = s_feb + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_c2b + 128);   // [0] weights landed, [1+g] operand ready, [3+g] accumulator ready
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows_static;
  const int n_tiles = (n_rows + TM - 1) / TM;

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
  if (tid < 128) {
    s_few[tid] = __ldg(a.w.fe_w + tid);
    s_feb[tid] = __ldg(a.w.fe_b + tid);
    s_c2b[tid] = __ldg(a.w.c2b + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  if (tmem != 0u) {   // the MMA issue uses compile-time TMEM addresses: the allocation of all 512 columns must start at column 0
    if (tid == 0) atomicOr(a.range_flag, 2);   // (cannot happen with one CTA per SM; if it ever does, the host re-runs in mode 1)
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
    return;
  }

  if (tid == 0 && static_cast<int>(blockIdx.x) * 2 < n_tiles) {   // all weight images, once
    mbar_expect_tx(&bars[0], (LOCAL ? 3 : 2) * IMG16_128);
    for (uint32_t off = 0; off < IMG16_128; off += 16384) {
      bulk_g2s(w1 + off, reinterpret_cast<const uint8_t*>(a.hW1) + off, 16384, &bars[0]);
      bulk_g2s(m2 + off, reinterpret_cast<const uint8_t*>(a.hM2) + off, 16384, &bars[0]);
      if (LOCAL) bulk_g2s(c2 + off, reinterpret_cast<const uint8_t*>(a.hC2) + off, 16384, &bars[0]);
    }
  }
  {
    const int g = warp >> 3, quad = warp & 3, half = (warp >> 2) & 1;
    const int my_row = quad * 32 + lane;
    Slot16 sl;
    sl.slot = tmem + static_cast<uint32_t>(g * SLOT_COLS);
    sl.trow = sl.slot + (static_cast<uint32_t>(quad * 32) << 16);
    sl.a_ready = &bars[1 + g];
    sl.d_ready = &bars[3 + g];
    sl.dph = 0;
    sl.aph = 0;
    sl.issuer = (warp & (F16_GWARPS - 1)) == 0;
    sl.scaled = a.scaled != 0;
    const float inv1 = __ldg(a.wsc + 0), inv2 = __ldg(a.wsc + 1), inv3 = __ldg(a.wsc + 2);
    const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    if (sl.issuer && static_cast<int>(blockIdx.x) * 2 + g < n_tiles) mbar_wait(&bars[0], 0);

    for (int it = 0;; ++it) {
      const int tile = (it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x)) * 2 + g;
      if (tile >= n_tiles) break;
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
          if (half == 0) {
            a.len_csc[r] = d;
            a.len_canon[__ldg(a.canon + r)] = d;
          }
        } else {
          d = __ldg(a.e_len + r);
        }
      }
      // ---- A = gelu(feature_expansion(d)): this thread's 64 input features
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float t[32];
        const int k0 = half * 64 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = gelu_fast(fmaf(s_few[k0 + j], d, s_feb[k0 + j]));
        store_split32(sl.trow, (k0 >> 1), t, lo_scale, amax);
      }
      sl.run_layer<HID, HID>(smem_u32(w1), IMG16_128 / 2);          // edge_feature_mlp.0 (x half)
      sl.wait_layer();
      // ---- g1 = gelu(D + T1[type]) -> A
      {
        const float* T1 = a.w.T1 + type * HID;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n0 = half * 64 + c * 32;
          uint32_t v[32];
          tmem_ld32(sl.trow + C16_D + n0, v);
          wait_ld();
          float t[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 tb = __ldg(reinterpret_cast<const float4*>(T1 + n0) + q);
            t[q * 4 + 0] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 0]), inv1, tb.x));
            t[q * 4 + 1] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 1]), inv1, tb.y));
            t[q * 4 + 2] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 2]), inv1, tb.z));
            t[q * 4 + 3] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 3]), inv1, tb.w));
          }
          store_split32(sl.trow, (n0 >> 1), t, lo_scale, amax);
        }
      }
      sl.run_layer<HID, HID>(smem_u32(m2), IMG16_128 / 2);          // combination_mlp.0 o edge_feature_mlp.2
      sl.wait_layer();
      // ---- g2 = gelu(D + T2[type])
      {
        const float* T2 = a.w.T2 + type * HID;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n0 = half * 64 + c * 32;
          uint32_t v[32];
          tmem_ld32(sl.trow + C16_D + n0, v);
          wait_ld();
          float t[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 tb = __ldg(reinterpret_cast<const float4*>(T2 + n0) + q);
            t[q * 4 + 0] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 0]), inv2, tb.x));
            t[q * 4 + 1] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 1]), inv2, tb.y));
            t[q * 4 + 2] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 2]), inv2, tb.z));
            t[q * 4 + 3] = gelu_fast(fmaf(__uint_as_float(v[q * 4 + 3]), inv2, tb.w));
          }
          if (LOCAL) {
            store_split32(sl.trow, (n0 >> 1), t, lo_scale, amax);
          } else {
            // the 32 features as packed fp16 hi / lo' pairs: words n0/2 .. n0/2+15 of the row and of its lo' half
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split2_f16(t[2 * j], t[2 * j + 1], lo_scale, hi[j], lo[j], amax);
            if (valid && a.g2) {
              float4* dst = reinterpret_cast<float4*>(a.g2 + r * HID + n0);
#pragma unroll
              for (int q = 0; q < 8; ++q) dst[q] = make_float4(t[q * 4], t[q * 4 + 1], t[q * 4 + 2], t[q * 4 + 3]);
            }
            if (valid) {
              const int w4 = n0 >> 3;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                a.g2h[g2h_index(r, w4 + q)] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                a.g2h[g2h_index(r, 16 + w4 + q)] = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
              }
            }
          }
        }
      }
      if (LOCAL) {
        sl.run_layer<HID, HID>(smem_u32(c2), IMG16_128 / 2);        // combination_mlp.2
        sl.wait_layer();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n0 = half * 64 + c * 32;
          uint32_t v[32];
          tmem_ld32(sl.trow + C16_D + n0, v);
          wait_ld();
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.out + r * HID + n0);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(fmaf(__uint_as_float(v[q * 4 + 0]), inv3, s_c2b[n0 + q * 4 + 0]),
                                   fmaf(__uint_as_float(v[q * 4 + 1]), inv3, s_c2b[n0 + q * 4 + 1]),
                                   fmaf(__uint_as_float(v[q * 4 + 2]), inv3, s_c2b[n0 + q * 4 + 2]),
                                   fmaf(__uint_as_float(v[q * 4 + 3]), inv3, s_c2b[n0 + q * 4 + 3]));
          }
        }
      }
      // D is only overwritten after this thread's next arrive on a_ready, which follows the wait::ld above in program order;
      // the A columns are dead since the last layer completed
    }
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int grid16(int64_t rows_cap, int num_sms) {
  int64_t pairs = (rows_cap + 2 * TM - 1) / (2 * TM);
  if (pairs < 1) pairs = 1;
  return (int)(pairs < num_sms ? pairs : num_sms);
}

void launch_encoder_global_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  TcEnc16Args a{};
  a.w = w.enc;
  a.hW1 = reinterpret_cast<const uint32_t*>(w.henc_W1); a.hM2 = reinterpret_cast<const uint32_t*>(w.henc_M2);
  a.hC2 = reinterpret_cast<const uint32_t*>(w.henc_C2); a.wsc = w.henc_sc;
  a.n_rows_dev = b.counters;
  a.e_len = b.e_len; a.e_type = b.e_type;
  a.g2h = b.g2h;
  a.g2 = c.f16_pair ? nullptr : b.g2;
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  tc_encoder16_kernel<false><<<grid16(b.cap, c.num_sms), F16_THREADS, TcEnc16Smem<false>::bytes, c.stream>>>(a);
  note_launch(c, "encoder.global_f16");
}

void launch_encoder_local_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos) {
  if (b.n_local == 0) return;
  TcEnc16Args a{};
  a.w = w.enc;
  a.hW1 = reinterpret_cast<const uint32_t*>(w.henc_W1); a.hM2 = reinterpret_cast<const uint32_t*>(w.henc_M2);
  a.hC2 = reinterpret_cast<const uint32_t*>(w.henc_C2); a.wsc = w.henc_sc;
  a.n_rows_dev = nullptr; a.n_rows_static = b.n_local;
  a.e_type = b.lc_type; a.pos = pos; a.src = b.lc_src; a.dst = b.lc_dst; a.canon = b.lc_canon;
  a.len_csc = b.lc_len; a.len_canon = b.lcc_len; a.len_in = b.lc_len_in; a.out = b.ea_loc;
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  tc_encoder16_kernel<true><<<grid16(b.n_local, c.num_sms), F16_THREADS, TcEnc16Smem<true>::bytes, c.stream>>>(a);
  note_launch(c, "encoder.local_f16");
}


// ------------------------------------------------------------------------------------------------ pair MLP
struct TcPair16Args {
  PairW w;                             // p1b, p2b, p3w, p3b
  const uint32_t *hP1h, *hP1e, *hP2;   // fp16 images: 128x128 (lo' scaled), 128x128 (lo unscaled, same weight scale as P1h), 64x128
  const float* wsc;                    // [0] inverse scale of P1h and P1e, [2] of P2
  const int* n_rows_dev;
  int n_rows_static;
  const float* h;                      // node features [N][128]
  const float* hmax;                   // [N] max_k |h[i][k]| (row_absmax_kernel): bounds |h_src * h_dst| for the per-row scale
  const uint4* g2h;                    // global: pre-split encoder state (lo' scaled by 2^S)
  const float* feat;                   // local: edge_attr [rows][128] fp32
  const int *src, *dst, *canon;
  float *s_csc, *s_canon;
  int scaled;
  int* range_flag;
};

constexpr size_t TC_PAIR16_SMEM = 1024 + 2 * IMG16_128 + IMG16_64 + (128 + 64 + 64 + 2 * 2 * 2 * 128) * sizeof(float) + 8 * sizeof(uint64_t) + 64;

template <bool LOCAL>
__global__ void __launch_bounds__(F16_THREADS, 1) tc_pair16_kernel(const TcPair16Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wP1h = base;
  uint8_t* wP1e = base + IMG16_128;
  uint8_t* wP2 = base + 2 * IMG16_128;
  float* s_p1b = reinterpret_cast<float*>(base + 2 * IMG16_128 + IMG16_64);
  float* s_p2b = s_p1b + 128;
  float* s_p3w = s_p2b + 64;
  float* s_part = s_p3w + 64;                                   // [2 groups][2 tile parities][2 halves][128] partial scores
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 2 * 2 * 128);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows_static;
  const int n_tiles = (n_rows + TM - 1) / TM;

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
  if (tid < 128) s_p1b[tid] = __ldg(a.w.p1b + tid);
  if (tid < 64) {
    s_p2b[tid] = __ldg(a.w.p2b + tid);
    s_p3w[tid] = __ldg(a.w.p3w + tid);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s_tmem;
  if (tmem != 0u) {   // the MMA issue uses compile-time TMEM addresses: the allocation of all 512 columns must start at column 0
    if (tid == 0) atomicOr(a.range_flag, 2);   // (cannot happen with one CTA per SM; if it ever does, the host re-runs in mode 1)
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
    return;
  }

  if (tid == 0 && static_cast<int>(blockIdx.x) * 2 < n_tiles) {   // all weight images, once
    mbar_expect_tx(&bars[0], 2 * IMG16_128 + IMG16_64);
    for (uint32_t off = 0; off < IMG16_128; off += 16384) {
      bulk_g2s(wP1h + off, reinterpret_cast<const uint8_t*>(a.hP1h) + off, 16384, &bars[0]);
      bulk_g2s(wP1e + off, reinterpret_cast<const uint8_t*>(a.hP1e) + off, 16384, &bars[0]);
    }
    for (uint32_t off = 0; off < IMG16_64; off += 16384) bulk_g2s(wP2 + off, reinterpret_cast<const uint8_t*>(a.hP2) + off, 16384, &bars[0]);
  }
  {
    const int g = warp >> 3, quad = warp & 3, half = (warp >> 2) & 1;
    const int my_row = quad * 32 + lane;
    Slot16 sl;
    sl.slot = tmem + static_cast<uint32_t>(g * SLOT_COLS);
    sl.trow = sl.slot + (static_cast<uint32_t>(quad * 32) << 16);
    sl.a_ready = &bars[1 + g];
    sl.d_ready = &bars[3 + g];
    sl.dph = 0;
    sl.aph = 0;
    sl.issuer = (warp & (F16_GWARPS - 1)) == 0;
    sl.scaled = a.scaled != 0;
    const float inv1 = __ldg(a.wsc + 0), inv2 = __ldg(a.wsc + 2);
    const float lo_scale = a.scaled ? static_cast<float>(1 << F16_LO_SHIFT) : 1.0f;
    const __half2 lo_unscale = __float2half2_rn(a.scaled ? 1.0f / static_cast<float>(1 << F16_LO_SHIFT) : 1.0f);
    const float p3b = __ldg(a.w.p3b);
    __half2 amax = __floats2half2_rn(0.f, 0.f);
    if (sl.issuer && static_cast<int>(blockIdx.x) * 2 + g < n_tiles) mbar_wait(&bars[0], 0);

    auto tile_of = [&](int it) { return (it * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x)) * 2 + g; };
    // h[src] and h[dst] slices of this thread's first 32 columns, one tile ahead (the other 32 columns: L1 prefetch)
    float4 hs[8], hd[8];
    int nsrc = 0, ndst = 0;
    float bound = 0.f;   // >= max |h_src * h_dst| of the row
    auto prefetch_h = [&](int tile) {
      const int64_t rr = static_cast<int64_t>(tile) * TM + my_row;
      if (tile < n_tiles && rr < n_rows) {
        nsrc = __ldg(a.src + rr);
        ndst = __ldg(a.dst + rr);
        bound = __ldg(a.hmax + nsrc) * __ldg(a.hmax + ndst);
        const float4* ps = reinterpret_cast<const float4*>(a.h + (size_t)nsrc * HID + half * 64);
        const float4* pd = reinterpret_cast<const float4*>(a.h + (size_t)ndst * HID + half * 64);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          hs[q] = __ldg(ps + q);
          hd[q] = __ldg(pd + q);
        }
        prefetch_l1(ps + 8);
        prefetch_l1(pd + 8);
      } else {
        nsrc = ndst = 0;
        bound = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) hs[q] = hd[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    prefetch_h(tile_of(0));

    for (int it = 0;; ++it) {
      const int tile = tile_of(it);
      if (tile >= n_tiles) break;
      const int64_t r = static_cast<int64_t>(tile) * TM + my_row;
      const bool valid = r < n_rows;
      // Per-row power-of-two scale rs = 2^-e that keeps the whole chain inside the fp16-split range: |h_src * h_dst| of random-
      // init (and possibly trained) networks reaches 1e5.  relu is positively homogeneous, so the scale rides through both
      // layers exactly (operands, biases) and is undone on the scalar score.  e = 0 (rs = 1: bit-identical arithmetic)
      // whenever the products stay below 2^11.
      float rs = 1.0f, rinv = 1.0f;
      if (bound > 2048.f && bound < 3.0e38f) {
        const int e = ilogbf(bound) - 10;
        rs = ldexpf(1.0f, -e);
        rinv = ldexpf(1.0f, e);
      }
      // ---- A = h[src] * h[dst]: 32 prefetched columns, then the other 32
      {
        float t[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          t[q * 4 + 0] = hs[q].x * rs * hd[q].x; t[q * 4 + 1] = hs[q].y * rs * hd[q].y;
          t[q * 4 + 2] = hs[q].z * rs * hd[q].z; t[q * 4 + 3] = hs[q].w * rs * hd[q].w;
        }
        store_split32(sl.trow, half * 32, t, lo_scale, amax);
        if (valid) {
          const float4* ps = reinterpret_cast<const float4*>(a.h + (size_t)nsrc * HID + half * 64 + 32);
          const float4* pd = reinterpret_cast<const float4*>(a.h + (size_t)ndst * HID + half * 64 + 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 x = __ldg(ps + q), y = __ldg(pd + q);
            t[q * 4 + 0] = x.x * rs * y.x; t[q * 4 + 1] = x.y * rs * y.y; t[q * 4 + 2] = x.z * rs * y.z; t[q * 4 + 3] = x.w * rs * y.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = 0.f;
        }
        store_split32(sl.trow, half * 32 + 16, t, lo_scale, amax);
      }
      sl.run_layer<HID, HID>(smem_u32(wP1h), IMG16_128 / 2);        // layers.0, h half
      // ---- the edge-feature rows of this tile travel while the tensor core works; their lo parts are UNSCALED (they
      //      accumulate onto the finished h half, where the scale-input-d fold is not available)
      uint32_t fhi[32], flo[32];
      if (LOCAL) {
        if (valid) {
          const float4* pf = reinterpret_cast<const float4*>(a.feat + r * HID + half * 64);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float4 v = __ldg(pf + q);
            split2_f16(v.x * rs, v.y * rs, 1.0f, fhi[2 * q], flo[2 * q], amax);
            split2_f16(v.z * rs, v.w * rs, 1.0f, fhi[2 * q + 1], flo[2 * q + 1], amax);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) fhi[j] = flo[j] = 0u;
        }
      } else {
        if (valid) {
          const __half2 rs2 = __float2half2_rn(rs);                              // exact (power of two >= 2^-24) ...
          const __half2 ls2 = __float2half2_rn(rs) * lo_unscale;                 // ... underflow only costs sub-2^-24 precision
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint4 vh = ldg_stream(a.g2h + g2h_index(r, 8 * half + q));
            const uint4 vl = ldg_stream(a.g2h + g2h_index(r, 16 + 8 * half + q));
            const uint32_t h4[4] = {vh.x, vh.y, vh.z, vh.w};
            const uint32_t l4[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half2 hv = (rs == 1.0f) ? *reinterpret_cast<const __half2*>(&h4[u]) : __hmul2(*reinterpret_cast<const __half2*>(&h4[u]), rs2);
              const __half2 lv = __hmul2(*reinterpret_cast<const __half2*>(&l4[u]), ls2);
              fhi[4 * q + u] = *reinterpret_cast<const uint32_t*>(&hv);
              flo[4 * q + u] = *reinterpret_cast<const uint32_t*>(&lv);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) fhi[j] = flo[j] = 0u;
        }
      }
      sl.wait_layer();
      // ---- A = edge features (the first product is complete, A may be overwritten; D keeps accumulating)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hi[j] = fhi[c * 16 + j];
          lo[j] = flo[c * 16 + j];
        }
        tmem_st16(sl.trow + C16_AHI + half * 32 + c * 16, hi);
        tmem_st16(sl.trow + C16_ALO + half * 32 + c * 16, lo);
      }
      {
        wait_st();
        fence_before_sync();
        mbar_arrive(sl.a_ready);
        if (sl.issuer) {
          mbar_wait(sl.a_ready, sl.aph);
          fence_after_sync();
          if (elect_one()) {
            if (sl.slot == 0u) issue_3xf16_acc_ct<HID, HID, 0u>(smem_u32(wP1e), IMG16_128 / 2);   // layers.0, edge half, accumulated
            else issue_3xf16_acc_ct<HID, HID, SLOT_COLS>(smem_u32(wP1e), IMG16_128 / 2);
            mma_commit(sl.d_ready);
          }
          __syncwarp();
        }
        sl.aph ^= 1u;
      }
      sl.wait_layer();
      // ---- r1 = relu(D + b1) -> A
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n0 = half * 64 + c * 32;
        uint32_t v[32];
        tmem_ld32(sl.trow + C16_D + n0, v);
        wait_ld();
        float t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = relu_(fmaf(__uint_as_float(v[j]), inv1, s_p1b[n0 + j] * rs));
        store_split32(sl.trow, (n0 >> 1), t, lo_scale, amax);
      }
      sl.run_layer<HID, 64>(smem_u32(wP2), IMG16_64 / 2);          // layers.1
      prefetch_h(tile_of(it + 1));                                   // next tile's node rows travel while layer 2 runs
      sl.wait_layer();
      // ---- score = layers.2(relu(D + b2)): 32 columns per thread, the two halves combined through smem
      float* part = s_part + (g * 2 + (it & 1)) * 256;
      {
        const int n0 = half * 32;
        uint32_t v[32];
        tmem_ld32(sl.trow + C16_D + n0, v);
        wait_ld();
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaf(relu_(fmaf(__uint_as_float(v[j]), inv2, s_p2b[n0 + j] * rs)), s_p3w[n0 + j], acc);
        part[half * 128 + my_row] = acc * rinv;
      }
      group_sync(1 + g, F16_GROUP);   // (double-buffered by tile parity: one barrier per tile is enough)
      if (half == 0 && valid) {
        const float s = (part[my_row] + part[128 + my_row]) + p3b;
        a.s_csc[r] = s;
        a.s_canon[__ldg(a.canon + r)] = s;
      }
    }
    if (f16_out_of_range(amax)) atomicOr(a.range_flag, 1);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// max_k |x[i][k]| per row of a [n][128] matrix, one warp per row
__global__ void __launch_bounds__(256) row_absmax_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * HID) + lane);
  float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
  // NaN-propagating maximum is not needed: a NaN row makes the bound NaN -> the comparisons fail -> scale 1, NaN flows on
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[row] = m;
}
static void launch_row_absmax(const LaunchCtx& c, const float* x, int n, float* out) {
  row_absmax_kernel<<<(n + 7) / 8, 256, 0, c.stream>>>(x, n, out);
  note_launch(c, "pair.row_absmax");
}

void launch_pair_global_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  launch_row_absmax(c, b.h, b.n_atoms, b.hmax);
  TcPair16Args a{};
  a.w = w.pg;
  a.hP1h = reinterpret_cast<const uint32_t*>(w.hpg_P1h); a.hP1e = reinterpret_cast<const uint32_t*>(w.hpg_P1e);
  a.hP2 = reinterpret_cast<const uint32_t*>(w.hpg_P2); a.wsc = w.hpg_sc;
  a.n_rows_dev = b.counters;
  a.h = b.h; a.hmax = b.hmax; a.g2h = b.g2h; a.src = b.e_src; a.dst = b.e_dst; a.canon = b.e_canon;
  a.s_csc = b.s_csc; a.s_canon = b.s_canon;
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  tc_pair16_kernel<false><<<grid16(b.cap, c.num_sms), F16_THREADS, TC_PAIR16_SMEM, c.stream>>>(a);
  note_launch(c, "pair.global_f16");
}

void launch_pair_local_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* h_local) {
  if (b.n_local == 0) return;
  launch_row_absmax(c, h_local, b.n_atoms, b.hmax);
  TcPair16Args a{};
  a.w = w.pl;
  a.hP1h = reinterpret_cast<const uint32_t*>(w.hpl_P1h); a.hP1e = reinterpret_cast<const uint32_t*>(w.hpl_P1e);
  a.hP2 = reinterpret_cast<const uint32_t*>(w.hpl_P2); a.wsc = w.hpl_sc;
  a.n_rows_dev = nullptr; a.n_rows_static = b.n_local;
  a.h = h_local; a.hmax = b.hmax; a.feat = b.ea_loc; a.src = b.lc_src; a.dst = b.lc_dst; a.canon = b.lc_canon;
  a.s_csc = b.sl_csc; a.s_canon = b.sl_canon;
  a.scaled = f16_lo_shift() != 0;
  a.range_flag = b.counters + 4;
  tc_pair16_kernel<true><<<grid16(b.n_local, c.num_sms), F16_THREADS, TC_PAIR16_SMEM, c.stream>>>(a);
  note_launch(c, "pair.local_f16");
}

void set_tc_mlp16_attributes() {
  cudaFuncSetAttribute(tc_pair16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_PAIR16_SMEM);
  cudaFuncSetAttribute(tc_pair16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_PAIR16_SMEM);
  cudaFuncSetAttribute(tc_encoder16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcEnc16Smem<false>::bytes);
  cudaFuncSetAttribute(tc_encoder16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcEnc16Smem<true>::bytes);
}

}  // namespace agd
