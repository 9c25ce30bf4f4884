// GIN local encoder (gin.py:38-69,112-148) over the static local edges (edge_type > 0).
// One kernel per layer, one CTA per 128 atoms: the message pass
//     m_i = sum_{e: dst_e = i} relu(x[src_e] + edge_attr_e) + (1 + eps) * x_i
// is a warp-per-atom gather over CSC segments straight into the transposed shared-memory tile,
// followed by the two 128x128 Linears (BatchNorm folded into the second), ReLU and the residual.
#include "common.cuh"
#include "kernels.h"

namespace agd {

constexpr size_t GIN_SMEM = (AS_FLOATS + WS_FLOATS) * sizeof(float);

__global__ void gin_embed_kernel(const float* __restrict__ emb, const int* __restrict__ atom_type, int n_nodes,
                                 float* __restrict__ x) {
  const int64_t total = (int64_t)n_nodes * (HID / 4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int node = (int)(i >> 5), q = (int)(i & 31);
    reinterpret_cast<float4*>(x)[i] = __ldg(reinterpret_cast<const float4*>(emb + (size_t)atom_type[node] * HID) + q);
  }
}

struct GinArgs {
  GinW w;
  const float* x_in;
  float* x_out;
  const float* ea;     // [n_local][128] edge_attr of the local edges, CSC order
  const int *src, *in_ptr;
  const int* ea_idx;   // row of `ea` for each CSC local edge (pair mode), or nullptr: the edge's own row
  int n_nodes;
  int last;
};

__global__ void __launch_bounds__(NT, 2) gin_layer_kernel(const GinArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Ws = As + AS_FLOATS;
  const TileCoord tc = tile_coord();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_rows = a.n_nodes;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const float ope = __ldg(a.w.sc);  // 1 + eps
  for (int mi = warp; mi < TM; mi += NT / 32) {
    const int64_t node = row0 + mi;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (node < n_rows) {
      const float* xi = a.x_in + (size_t)node * HID;
      float self[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) self[j] = __ldg(xi + lane + 32 * j);
      const int e0 = a.in_ptr[node], e1 = a.in_ptr[node + 1];
      for (int e = e0; e < e1; ++e) {
        const float* xs = a.x_in + (size_t)__ldg(a.src + e) * HID;
        const float* ee = a.ea + (size_t)(a.ea_idx ? __ldg(a.ea_idx + e) : e) * HID;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += relu_(__ldg(xs + lane + 32 * j) + __ldg(ee + lane + 32 * j));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaf(ope, self[j], v[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) As[(lane + 32 * j) * LDA + mi] = v[j];
  }
  float acc[8][8];
  tile_gemm<HID, HID, false>(a.w.G1, As, Ws, acc, tc.tx, tc.ty);
  tile_store_smem<HID>(acc, As, tc.tx, tc.ty, [&](float v, int m, int n) { return relu_(v + __ldg(a.w.g1b + n)); });
  tile_gemm<HID, HID, false>(a.w.G2, As, Ws, acc, tc.tx, tc.ty);
  const int last = a.last;
  tile_store_global<HID>(acc, a.x_out, row0, n_rows, HID, 0, tc.tx, tc.ty, [&](float v, int m, int n) {
    float o = v + __ldg(a.w.g2b + n);
    if (!last) o = relu_(o);
    return o + __ldg(a.x_in + (size_t)(row0 + m) * HID + n);
  });
}

void launch_gin_embed(const LaunchCtx& c, const BatchDev& b, const ModelW& w) {
  int64_t blocks = ((int64_t)b.n_atoms * 32 + 255) / 256;
  if (blocks > c.num_sms * 8) blocks = c.num_sms * 8;
  if (blocks < 1) blocks = 1;
  gin_embed_kernel<<<(int)blocks, 256, 0, c.stream>>>(w.gin_emb, b.atom_type, b.n_atoms, b.gx0);
  note_launch(c, "gin.embed");
}

void launch_gin_layer(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int layer, const float* x_in, float* x_out) {
  GinArgs a{};
  a.w = w.gin[layer];
  a.x_in = x_in;
  a.x_out = x_out;
  a.ea = b.ea_loc;
  a.src = b.lc_src;
  a.in_ptr = b.lc_in_ptr;
  a.ea_idx = b.lc_ea_idx;
  a.n_nodes = b.n_atoms;
  a.last = (layer == c.num_convs_local - 1) ? 1 : 0;
  gin_layer_kernel<<<(b.n_atoms + TM - 1) / TM, NT, GIN_SMEM, c.stream>>>(a);
  note_launch(c, "gin.layer");
}

void set_gin_attributes() {
  cudaFuncSetAttribute(gin_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GIN_SMEM);
}

}  // namespace agd
