// Per-edge MLP chains: the MLP edge encoder (edge.py:84-103) and the pair MLPs
// (common.py:86-109 via dualenc.py:203-211,226-235).  Each CTA owns a tile of 128 edges and runs
// the whole chain with activations resident in shared memory (see common.cuh).
//
// Algebra done once on the host (pack.py), so that no FLOP is spent on it per edge:
//   * the bond-embedding half of both 256->128 Linears is a per-type table (T1, T2);
//   * edge_feature_mlp.2 followed by combination_mlp.0 is one matrix M2;
//   * the attention branch multiplies by softmax over a size-1 dim == 1.0 exactly, so it is skipped;
//   * on the global branch edge_attr is only ever consumed by Linears (CFConv filter nets, pair
//     MLP), so combination_mlp.2 is merged into those and the kernel emits g2 = gelu(...) instead.
#include "common.cuh"
#include "kernels.h"

namespace agd {

constexpr size_t ENC_SMEM = (AS_FLOATS + WS_FLOATS) * sizeof(float) + TM * (sizeof(int) + sizeof(float));
constexpr size_t PAIR_SMEM = (AS_FLOATS + WS_FLOATS) * sizeof(float) + TM * 2 * sizeof(int) + 2 * TM * sizeof(float);

struct EncArgs {
  EncW w;
  const int* n_rows_dev;   // device row count (global) or nullptr
  int n_rows_static;       // used when n_rows_dev == nullptr
  int max_tiles;
  // global: precomputed per-edge inputs
  const float* e_len;
  const int* e_type;
  // local: lengths from positions
  const float* pos;
  const int *src, *dst, *canon;
  float *len_csc, *len_canon;
  const float* len_in;     // local: caller-supplied lengths instead of |pos[src]-pos[dst]|
  float* out;              // g2 (global) or edge_attr (local), [rows][128]
};

template <bool LOCAL>
__global__ void __launch_bounds__(NT, 2) edge_encoder_kernel(const EncArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Ws = As + AS_FLOATS;
  int* s_type = reinterpret_cast<int*>(Ws + WS_FLOATS);
  float* s_len = reinterpret_cast<float*>(s_type + TM);
  const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows_static;
  const int n_tiles = (n_rows + TM - 1) / TM;
  const TileCoord tc = tile_coord();
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();  // previous tile's epilogue reads of s_type are done
    if (tid < TM) {
      const int64_t r = row0 + tid;
      int t = 0;
      float d = 0.f;
      if (r < n_rows) {
        if (LOCAL) {
          const int s = a.src[r], q = a.dst[r];
          const float dx = a.pos[3 * (size_t)s] - a.pos[3 * (size_t)q];
          const float dy = a.pos[3 * (size_t)s + 1] - a.pos[3 * (size_t)q + 1];
          const float dz = a.pos[3 * (size_t)s + 2] - a.pos[3 * (size_t)q + 2];
          d = a.len_in ? a.len_in[r] : sqrtf(dx * dx + dy * dy + dz * dz);
          a.len_csc[r] = d;
          a.len_canon[a.canon[r]] = d;
        } else {
          d = a.e_len[r];
        }
        t = a.e_type[r];
      }
      s_type[tid] = t;
      s_len[tid] = d;
    }
    __syncthreads();
    // x = gelu(feature_expansion(d))  -> As[k][m]
    {
      const int m = tid & (TM - 1);
      const float d = s_len[m];
#pragma unroll 4
      for (int k = tid >> 7; k < HID; k += 2) As[k * LDA + m] = gelu_erf(fmaf(__ldg(a.w.fe_w + k), d, __ldg(a.w.fe_b + k)));
    }
    float acc[8][8];
    tile_gemm<HID, HID, false>(a.w.W1, As, Ws, acc, tc.tx, tc.ty);
    tile_store_smem<HID>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) {
      return gelu_erf(v + __ldg(a.w.T1 + s_type[m] * HID + n));
    });
    tile_gemm<HID, HID, false>(a.w.M2, As, Ws, acc, tc.tx, tc.ty);
    if (!LOCAL) {
      tile_store_global<HID>(acc, a.out, row0, n_rows, HID, 0, tc.tx, tc.ty, [&](float v, int m, int n) {
        return gelu_erf(v + __ldg(a.w.T2 + s_type[m] * HID + n));
      });
    } else {
      tile_store_smem<HID>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) {
        return gelu_erf(v + __ldg(a.w.T2 + s_type[m] * HID + n));
      });
      tile_gemm<HID, HID, false>(a.w.C2, As, Ws, acc, tc.tx, tc.ty);
      tile_store_global<HID>(acc, a.out, row0, n_rows, HID, 0, tc.tx, tc.ty,
                             [&](float v, int m, int n) { return v + __ldg(a.w.c2b + n); });
    }
  }
}

struct PairArgs {
  PairW w;
  const int* n_rows_dev;
  int n_rows_static;
  int max_tiles;
  const float* h;       // node features [N][128]
  const float* feat;    // g2 (global) / edge_attr (local) [rows][128]
  const int *src, *dst, *canon;
  float *s_csc, *s_canon;
  int act;              // mlp_act (common.py:44-84 takes any torch.nn.functional name): AGD_ACT_* of kernels.h
};

// mlp_act of the pair MLPs (MultiLayerPerceptron, common.py:59-62: getattr(F, activation)), torch's fp32 definitions
__device__ __forceinline__ float pair_act(float x, int act) {
  switch (act) {
    case AGD_ACT_GELU: return gelu_erf(x);
    case AGD_ACT_SILU: return x / (1.0f + expf(-x));
    case AGD_ACT_TANH: return tanhf(x);
    case AGD_ACT_SIGMOID: return sigmoidf_(x);
    case AGD_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
    case AGD_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case AGD_ACT_SOFTPLUS: return x > 20.0f ? x : log1pf(expf(x));
    default: return relu_(x);
  }
}

// edge_inv = MLP([h[row]*h[col], edge_attr]); first Linear split in two K=128 passes so one
// activation tile suffices (2 CTAs/SM).
__global__ void __launch_bounds__(NT, 2) pair_mlp_kernel(const PairArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Ws = As + AS_FLOATS;
  int* s_src = reinterpret_cast<int*>(Ws + WS_FLOATS);
  int* s_dst = s_src + TM;
  float* s_red = reinterpret_cast<float*>(s_dst + TM);  // [2][TM]
  const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows_static;
  const int n_tiles = (n_rows + TM - 1) / TM;
  const TileCoord tc = tile_coord();
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    if (tid < TM) {
      const int64_t r = row0 + tid;
      s_src[tid] = (r < n_rows) ? a.src[r] : 0;
      s_dst[tid] = (r < n_rows) ? a.dst[r] : 0;
    }
    __syncthreads();
    tile_load_pair_T(a.h, s_src, s_dst, As);
    float acc[8][8];
    tile_gemm<HID, HID, false>(a.w.P1h, As, Ws, acc, tc.tx, tc.ty);
    tile_load_T<HID>(a.feat, row0, n_rows, HID, 0, As);
    tile_gemm<HID, HID, true>(a.w.P1e, As, Ws, acc, tc.tx, tc.ty);
    tile_store_smem<HID>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) { return pair_act(v + __ldg(a.w.p1b + n), a.act); });
    float acc2[8][4];
    tile_gemm<HID, 64, false>(a.w.P2, As, Ws, acc2, tc.tx, tc.ty);
    // last Linear (64 -> 1): per-thread partial over its 4 columns, reduce over the 16 tx threads
    float part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = tc.tx * 4 + j;
      const float wj = __ldg(a.w.p3w + n), bj = __ldg(a.w.p2b + n);
#pragma unroll
      for (int i = 0; i < 8; ++i) part[i] = fmaf(pair_act(acc2[i][j] + bj, a.act), wj, part[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1);
      part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2);
      part[i] += __shfl_xor_sync(0xffffffffu, part[i], 4);
    }
    if ((tid & 7) == 0) {
      const int half = (tid >> 5) & 1;  // which group of 8 tx values this warp holds
#pragma unroll
      for (int i = 0; i < 8; ++i) s_red[half * TM + tile_row(tc.ty, i)] = part[i];
    }
    __syncthreads();
    if (tid < TM) {
      const int64_t r = row0 + tid;
      if (r < n_rows) {
        const float s = s_red[tid] + s_red[TM + tid] + __ldg(a.w.p3b);
        a.s_csc[r] = s;
        a.s_canon[a.canon[r]] = s;
      }
    }
  }
}

static int tiles_grid(int64_t rows_cap, int num_sms, int per_sm) {
  int64_t t = (rows_cap + TM - 1) / TM;
  if (t < 1) t = 1;
  const int64_t g = (int64_t)num_sms * per_sm;
  return (int)(t < g ? t : g);
}

void launch_encoder_global(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  EncArgs a{};
  a.w = w.enc;
  a.n_rows_dev = b.counters;
  a.e_len = b.e_len;
  a.e_type = b.e_type;
  a.out = b.g2;
  edge_encoder_kernel<false><<<tiles_grid(b.cap, c.num_sms, 2), NT, ENC_SMEM, c.stream>>>(a);
  note_launch(c, "encoder.global");
}

void launch_encoder_local(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos) {
  if (b.n_local == 0) return;
  EncArgs a{};
  a.w = w.enc;
  a.n_rows_dev = nullptr;
  a.n_rows_static = b.n_local;
  a.e_type = b.lc_type;
  a.pos = pos;
  a.src = b.lc_src;
  a.dst = b.lc_dst;
  a.canon = b.lc_canon;
  a.len_csc = b.lc_len;
  a.len_canon = b.lcc_len;
  a.len_in = b.lc_len_in;
  a.out = b.ea_loc;
  edge_encoder_kernel<true><<<tiles_grid(b.n_local, c.num_sms, 2), NT, ENC_SMEM, c.stream>>>(a);
  note_launch(c, "encoder.local");
}

void launch_pair_global(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  PairArgs a{};
  a.w = w.pg;
  a.n_rows_dev = b.counters;
  a.h = b.h;
  a.feat = b.g2;
  a.src = b.e_src;
  a.dst = b.e_dst;
  a.canon = b.e_canon;
  a.s_csc = b.s_csc;
  a.s_canon = b.s_canon;
  a.act = c.mlp_act;
  pair_mlp_kernel<<<tiles_grid(b.cap, c.num_sms, 2), NT, PAIR_SMEM, c.stream>>>(a);
  note_launch(c, "pair.global");
}

void launch_pair_local(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* h_local) {
  if (b.n_local == 0) return;
  PairArgs a{};
  a.w = w.pl;
  a.n_rows_dev = nullptr;
  a.n_rows_static = b.n_local;
  a.h = h_local;
  a.feat = b.ea_loc;
  a.src = b.lc_src;
  a.dst = b.lc_dst;
  a.canon = b.lc_canon;
  a.s_csc = b.sl_csc;
  a.s_canon = b.sl_canon;
  a.act = c.mlp_act;
  pair_mlp_kernel<<<tiles_grid(b.n_local, c.num_sms, 2), NT, PAIR_SMEM, c.stream>>>(a);
  note_launch(c, "pair.local");
}

void set_encoder_attributes() {
  cudaFuncSetAttribute(edge_encoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM);
  cudaFuncSetAttribute(edge_encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_SMEM);
  cudaFuncSetAttribute(pair_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM);
}

}  // namespace agd
