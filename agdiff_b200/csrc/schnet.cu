// Attention-augmented SchNet global encoder (schnet.py:113-282) as three kernels per block:
//   filters    per-edge filter-generating nets of conv1 (F=128) and conv2 (F=64):
//              W = (Lin(SSP_beta(Lin'(g2)))) * sigmoid(distance MLP) * cutoff envelope  -> filt[E][192]
//   aggregate  agg[i] = sum over in-edges e of x[src_e] (.) W_e   (CSC segments, no atomics, deterministic)
//   node       lin2+BN | concat | SSP | lin | attention gate | adaptive scaling | residual |
//              next block's lin1+BN+LeakyReLU                                            -> h, xcat
#include "common.cuh"
#include "kernels.h"

namespace agd {

constexpr size_t FILT_SMEM = (AS_FLOATS + WS_FLOATS) * sizeof(float) + TM * sizeof(float);
constexpr size_t NODE_SMEM = (2 * AS_FLOATS + WS_FLOATS) * sizeof(float) + (TM * 2 + 8 * TM) * sizeof(float) + TM * sizeof(int);

struct FiltArgs {
  const float *F1, *f1b, *F2, *f2b, *dw;
  const float* beta_ptr;
  const int* n_rows_dev;
  const float* g2;
  const float* e_len;
  float* filt;      // [rows][192]
  int col0;         // 0 (conv1) or 128 (conv2)
  float cutoff;
  int smooth;
};

template <int F>
__global__ void __launch_bounds__(NT, 2) filter_kernel(const FiltArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Ws = As + AS_FLOATS;
  float* s_cw = Ws + WS_FLOATS;
  const int n_rows = *a.n_rows_dev;
  const int n_tiles = (n_rows + TM - 1) / TM;
  const TileCoord tc = tile_coord();
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    if (tid < TM) {
      const int64_t r = row0 + tid;
      s_cw[tid] = (r < n_rows) ? cfconv_edge_weight(a.e_len[r], a.dw, a.cutoff, a.smooth) : 0.f;
    }
    tile_load_T<HID>(a.g2, row0, n_rows, HID, 0, As);
    float acc[8][F / 16];
    tile_gemm<HID, F, false>(a.F1, As, Ws, acc, tc.tx, tc.ty);
    const float beta = __ldg(a.beta_ptr);
    tile_store_smem<F>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) { return ssp(v + __ldg(a.f1b + n), beta); });
    tile_gemm<F, F, false>(a.F2, As, Ws, acc, tc.tx, tc.ty);
    tile_store_global<F>(acc, a.filt, row0, n_rows, 192, a.col0, tc.tx, tc.ty,
                         [&](float v, int m, int n) { return (v + __ldg(a.f2b + n)) * s_cw[m]; });
  }
}

// ------------------------------------------------------------------ gather -> multiply -> segmented sum
// One thread per (destination atom, float4 column); F/4 threads cover an atom, so the filter rows of
// its in-edges are read as one contiguous stream and the source rows as F*4-byte gathers (L2 hits:
// a molecule's x fits in a few KB).  HBM-bound: algorithmic bytes per edge = 4F (filter) + 4F (gather)
// + 4 (index), per atom 4F (write) + 4 (pointer).
// SUMMATION ORDER (the definition every CFConv aggregation in this library follows, tc_cfconv.cu included, so that they agree
// bit for bit): the in-edges of a destination, in CSC order, are dealt round-robin to FOUR partial sums by their position in
// the run (edge e0 + p goes to partial p mod 4), each partial accumulates its edges in order with one fmaf per edge, and the
// result is (s0 + s1) + (s2 + s3).  The position in the run does not depend on tile, CTA or batch boundaries, so neither does
// the result; four independent chains also give the gather four times the memory-level parallelism of one.
template <int F>
__global__ void __launch_bounds__(F) cfconv_aggregate_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                             const int* __restrict__ src, const int* __restrict__ in_ptr,
                                                             int n_nodes, float* __restrict__ out) {
  constexpr int TPN = F / 4;            // threads per node
  constexpr int NPB = F / TPN;          // nodes per block (= 4)
  const int node = blockIdx.x * NPB + threadIdx.x / TPN;
  const int c4 = threadIdx.x % TPN;
  if (node >= n_nodes) return;
  const int e0 = in_ptr[node], e1 = in_ptr[node + 1];
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = e0; e < e1; e += 4) {
    int s[4];
    float4 wv[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = (e + u < e1) ? __ldg(src + e + u) : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (e + u < e1) {
        wv[u] = __ldcs(reinterpret_cast<const float4*>(W + (size_t)(e + u) * F) + c4);
        xv[u] = __ldg(reinterpret_cast<const float4*>(x + (size_t)s[u] * F) + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (e + u < e1) {
        acc[u].x = fmaf(xv[u].x, wv[u].x, acc[u].x);
        acc[u].y = fmaf(xv[u].y, wv[u].y, acc[u].y);
        acc[u].z = fmaf(xv[u].z, wv[u].z, acc[u].z);
        acc[u].w = fmaf(xv[u].w, wv[u].w, acc[u].w);
      }
    }
  }
  float4 r;
  r.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x);
  r.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
  r.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z);
  r.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
  reinterpret_cast<float4*>(out + (size_t)node * F)[c4] = r;
}

// ------------------------------------------------------------------ node-side kernel
struct NodeArgs {
  BlkW w;            // block whose convs just aggregated (unused when first)
  const float *nL1a, *nl1ab, *nL1b, *nl1bb;   // NEXT block's lin1 (nullptr after the last block)
  const float* emb;  // SchNet embedding table (first only)
  const int* atom_type;
  int n_nodes;
  int first;         // 1: h = emb[z], then only the lin1 stage
  const float* agg;  // [N][192]
  float* h;          // [N][128] in/out
  float* xcat;       // [N][192] out
};

__global__ void __launch_bounds__(NT, 1) schnet_node_kernel(const NodeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + AS_FLOATS;
  float* Ws = Bs + AS_FLOATS;
  float* s_att = Ws + WS_FLOATS;       // [2][TM] attention logit halves
  float* s_r8 = s_att + 2 * TM;        // [8][TM]
  int* s_z = reinterpret_cast<int*>(s_r8 + 8 * TM);
  const TileCoord tc = tile_coord();
  const int tid = threadIdx.x;
  const int64_t n_rows = a.n_nodes;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  float acc[8][8];

  if (a.first) {
    if (tid < TM) s_z[tid] = (row0 + tid < n_rows) ? a.atom_type[row0 + tid] : 0;
    __syncthreads();
    // h = embedding[z] -> global h and As[k][m]
    for (int it = 0; it < TM * (HID / 4) / NT; ++it) {
      const int idx = it * NT + tid;
      const int m = idx & (TM - 1), kq = idx >> 7;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.emb + (size_t)s_z[m] * HID) + kq);
      As[(kq * 4 + 0) * LDA + m] = v.x;
      As[(kq * 4 + 1) * LDA + m] = v.y;
      As[(kq * 4 + 2) * LDA + m] = v.z;
      As[(kq * 4 + 3) * LDA + m] = v.w;
      if (row0 + m < n_rows) reinterpret_cast<float4*>(a.h + (size_t)(row0 + m) * HID)[kq] = v;
    }
  } else {
    const float beta_act = __ldg(a.w.sc + 2);
    // v1 = BN(lin2_1(agg[:, :128])), v2 = BN(lin2_2(agg[:, 128:])) ; t = SSP(cat[v1, v2])
    tile_load_T<HID>(a.agg, row0, n_rows, 192, 0, As);
    tile_load_T<64>(a.agg, row0, n_rows, 192, 128, Bs);
    tile_gemm<HID, HID, false>(a.w.L2a, As, Ws, acc, tc.tx, tc.ty);
    tile_store_smem<HID>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) { return ssp(v + __ldg(a.w.l2ab + n), beta_act); });
    tile_gemm<64, HID, false>(a.w.L2b, Bs, Ws, acc, tc.tx, tc.ty);
    tile_store_smem<HID>(acc, Bs, tc.tx, tc.ty, [&](float v, int m, int n) { return ssp(v + __ldg(a.w.l2bb + n), beta_act); });
    // xc = lin(t)
    tile_gemm<HID, HID, false>(a.w.LIN, As, Ws, acc, tc.tx, tc.ty);
    tile_gemm<HID, HID, true>(a.w.LIN + HID * HID, Bs, Ws, acc, tc.tx, tc.ty);
    tile_store_smem<HID>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) { return v + __ldg(a.w.linb + n); });
    // attention gate: sigmoid(a2 . relu(A1 xc + a1b) + a2b)
    {
      float acc2[8][4];
      tile_gemm<HID, 64, false>(a.w.A1, As, Ws, acc2, tc.tx, tc.ty);
      float part[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) part[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = tc.tx * 4 + j;
        const float wj = __ldg(a.w.a2w + n), bj = __ldg(a.w.a1b + n);
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i] = fmaf(relu_(acc2[i][j] + bj), wj, part[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1);
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2);
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 4);
      }
      if ((tid & 7) == 0) {
        const int half = (tid >> 5) & 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) s_att[half * TM + tile_row(tc.ty, i)] = part[i];
      }
    }
    __syncthreads();
    // y = xc * gate  (in place, lanes walk rows)
    {
      const int m = tid & (TM - 1);
      const float gate = sigmoidf_(s_att[m] + s_att[TM + m] + __ldg(a.w.sc + 3));
      for (int k = tid >> 7; k < HID; k += 2) As[k * LDA + m] *= gate;
    }
    __syncthreads();
    // adaptive scaling: r8 = relu(S1^T y) (128 -> 8), s = sigmoid(S2^T r8) (8 -> 128), out = y * s
    for (int idx = tid; idx < 8 * TM; idx += NT) {
      const int m = idx & (TM - 1), j = idx >> 7;
      float s = 0.f;
#pragma unroll 8
      for (int k = 0; k < HID; ++k) s = fmaf(As[k * LDA + m], __ldg(a.w.S1 + k * 8 + j), s);
      s_r8[j * TM + m] = relu_(s);
    }
    __syncthreads();
    {
      const int m = tid & (TM - 1);
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = s_r8[j * TM + m];
      for (int k = tid >> 7; k < HID; k += 2) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(r[j], __ldg(a.w.S2 + j * HID + k), s);
        As[k * LDA + m] *= sigmoidf_(s);
      }
    }
    __syncthreads();
    // h += out : coalesced read-modify-write of global h, result also back into As for lin1
    for (int it = 0; it < TM * (HID / 4) / NT; ++it) {
      const int idx = it * NT + tid;
      const int kq = idx & 31, m = idx >> 5;   // a warp covers one row
      if (row0 + m < n_rows) {
        float4* hp = reinterpret_cast<float4*>(a.h + (size_t)(row0 + m) * HID) + kq;
        float4 v = *hp;
        v.x += As[(kq * 4 + 0) * LDA + m];
        v.y += As[(kq * 4 + 1) * LDA + m];
        v.z += As[(kq * 4 + 2) * LDA + m];
        v.w += As[(kq * 4 + 3) * LDA + m];
        *hp = v;
        As[(kq * 4 + 0) * LDA + m] = v.x;
        As[(kq * 4 + 1) * LDA + m] = v.y;
        As[(kq * 4 + 2) * LDA + m] = v.z;
        As[(kq * 4 + 3) * LDA + m] = v.w;
      }
    }
  }
  if (a.nL1a == nullptr) return;
  // next block's x = LeakyReLU_0.2(BN(lin1(h))) for conv1 (128) and conv2 (64)
  tile_gemm<HID, HID, false>(a.nL1a, As, Ws, acc, tc.tx, tc.ty);
  tile_store_global<HID>(acc, a.xcat, row0, n_rows, 192, 0, tc.tx, tc.ty,
                         [&](float v, int m, int n) { return leaky02(v + __ldg(a.nl1ab + n)); });
  {
    float acc2[8][4];
    tile_gemm<HID, 64, false>(a.nL1b, As, Ws, acc2, tc.tx, tc.ty);
    tile_store_global<64>(acc2, a.xcat, row0, n_rows, 192, 128, tc.tx, tc.ty,
                          [&](float v, int m, int n) { return leaky02(v + __ldg(a.nl1bb + n)); });
  }
}

static int tiles_grid(int64_t rows_cap, int num_sms, int per_sm) {
  int64_t t = (rows_cap + TM - 1) / TM;
  if (t < 1) t = 1;
  const int64_t g = (int64_t)num_sms * per_sm;
  return (int)(t < g ? t : g);
}

void launch_filters(const LaunchCtx& c, const BatchDev& b, const ModelW& mw, int blk) {
  if (c.use_tc == 2) {
    if (c.f16_fuse) launch_cfconv_f16(c, b, mw, blk);
    else launch_filters_f16(c, b, mw, blk);
    return;
  }
  if (c.use_tc) {
    launch_filters_tc(c, b, mw, blk);
    return;
  }
  const BlkW& w = mw.blk[blk];
  FiltArgs a{};
  a.n_rows_dev = b.counters;
  a.g2 = b.g2;
  a.e_len = b.e_len;
  a.filt = b.filt;
  a.cutoff = c.cutoff;
  a.smooth = c.smooth;
  a.F1 = w.F1a; a.f1b = w.f1ab; a.F2 = w.F2a; a.f2b = w.f2ab; a.dw = w.dw1; a.beta_ptr = w.sc + 0; a.col0 = 0;
  filter_kernel<128><<<tiles_grid(b.cap, c.num_sms, 2), NT, FILT_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.filter128");
  a.F1 = w.F1b; a.f1b = w.f1bb; a.F2 = w.F2b; a.f2b = w.f2bb; a.dw = w.dw2; a.beta_ptr = w.sc + 1; a.col0 = 128;
  filter_kernel<64><<<tiles_grid(b.cap, c.num_sms, 2), NT, FILT_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.filter64");
}

void launch_aggregate(const LaunchCtx& c, const float* x, const float* W, const int* src, const int* in_ptr, int n_nodes,
                      int F, float* out) {
  if (n_nodes <= 0) return;
  const int blocks = (n_nodes + 3) / 4;
  if (F == 192)
    cfconv_aggregate_kernel<192><<<blocks, 192, 0, c.stream>>>(x, W, src, in_ptr, n_nodes, out);
  else if (F == 128)
    cfconv_aggregate_kernel<128><<<blocks, 128, 0, c.stream>>>(x, W, src, in_ptr, n_nodes, out);
  else
    cfconv_aggregate_kernel<64><<<blocks, 64, 0, c.stream>>>(x, W, src, in_ptr, n_nodes, out);
  note_launch(c, "schnet.aggregate");
}

void launch_schnet_node(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk) {
  NodeArgs a{};
  a.n_nodes = b.n_atoms;
  a.atom_type = b.atom_type;
  a.emb = w.sch_emb;
  a.agg = b.agg;
  a.h = b.h;
  a.xcat = b.xcat;
  a.first = (blk < 0) ? 1 : 0;
  if (blk >= 0) a.w = w.blk[blk];
  const int nxt = blk + 1;
  if (nxt < c.num_convs) {
    a.nL1a = w.blk[nxt].L1a; a.nl1ab = w.blk[nxt].l1ab; a.nL1b = w.blk[nxt].L1b; a.nl1bb = w.blk[nxt].l1bb;
  }
  const int blocks = (b.n_atoms + TM - 1) / TM;
  schnet_node_kernel<<<blocks, NT, NODE_SMEM, c.stream>>>(a);
  note_launch(c, "schnet.node");
}

void set_schnet_attributes() {
  cudaFuncSetAttribute(filter_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILT_SMEM);
  cudaFuncSetAttribute(filter_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FILT_SMEM);
  cudaFuncSetAttribute(schnet_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NODE_SMEM);
}

}  // namespace agd
