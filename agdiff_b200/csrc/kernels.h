// Host-side declarations shared by the translation units of libagdiff_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/agdiff_b200.h"

namespace agd {

enum { AGD_ACT_RELU = 0, AGD_ACT_GELU, AGD_ACT_SILU, AGD_ACT_TANH, AGD_ACT_SIGMOID, AGD_ACT_LEAKY_RELU, AGD_ACT_ELU, AGD_ACT_SOFTPLUS, AGD_ACT_COUNT };
constexpr int MAX_BLOCKS = 8;   // SchNet interaction blocks supported (reference config: 6)
constexpr int MAX_GIN = 8;      // GIN layers supported (reference config: 4)

// ------------------------------------------------------------------ packed weights (device pointers)
// All matrices are [K][N] row-major (= torch weight transposed), BatchNorm(eval) folded, and
// back-to-back Linears merged by the host mirror (agdiff_b200/pack.py) in fp64.
struct EncW {              // MLPEdgeEncoder edge.py:84-103 (the *global* instance serves both branches)
  const float *fe_w, *fe_b;   // [128] feature_expansion
  const float *T1, *W1;       // [100][128] per-type table (bond half of edge_feature_mlp.0 + bias), [128][128]
  const float *T2, *M2;       // [100][128] table, [128][128] = combination_mlp.0[:, :128] @ edge_feature_mlp.2
  const float *C2, *c2b;      // combination_mlp.2 (only the local branch materialises edge_attr)
};
struct BlkW {              // InteractionBlock k, schnet.py:165-216 (+ scaling module :219-234)
  const float *L1a, *l1ab, *L1b, *l1bb;      // conv{1,2}.lin1 with norm1 folded: [128][128], [128][64]
  const float *F1a, *f1ab, *F2a, *f2ab;      // conv1.nn: first layer merged with the encoder's last Linear
  const float *F1b, *f1bb, *F2b, *f2bb;      // conv2.nn: [128][64], [64][64]
  const float *dw1, *dw2;                    // distance_weighting packs [w1|b1|w2|b2] (100 floats used)
  const float *L2a, *l2ab, *L2b, *l2bb;      // conv{1,2}.lin2 with norm2 folded: [128][128], [64][128]
  const float *LIN, *linb;                   // lin: [256][128]
  const float *A1, *a1b, *a2w;               // attention: [128][64], [64], [64]
  const float *S1, *S2;                      // scaling fc: [128][8], [8][128]
  const float *sc;                           // scalars: beta(conv1.nn.1), beta(conv2.nn.1), beta(act), attention.2.bias
  const float *tF1a, *tF2a, *tF1b, *tF2b;    // tcgen05 operand images of the filter nets: [hi | lo], K-major SWIZZLE_128B
  const float *tL1a, *tL1b, *tL2a, *tL2b, *tLINa, *tLINb, *tA1;   // ... of the node-side Linears
  const float *hF1a, *hF2a, *hF1b, *hF2b;    // fp16-split operand images of the filter nets: [hi | lo'] packed halves (tc_filter16.cu)
  const float* hsc;                          // inverse power-of-two weight scales of hF1a, hF2a, hF1b, hF2b
  const float *hL2a, *hLINa, *hL2b, *hLINb, *hA1, *hL1a, *hL1b;   // fp16-split images of the node-side Linears (tc_node16.cu)
  const float* hnsc;                         // their inverse scales [L2a, LINa, L2b, LINb, A1, L1a, L1b, -]
};
struct PairW {             // grad_{global,local}_dist_mlp, common.py:86-103 on [h_row*h_col, edge_attr]
  const float *P1h, *P1e, *p1b;   // layers.0 split: [128][128] on h_row*h_col, [128][128] on g2 (global, merged) / edge_attr (local)
  const float *P2, *p2b;          // [128][64]
  const float *p3w, *p3b;         // [64], [1]
};
struct GinW {              // GINEConv + BatchNorm, gin.py:38-69,112-148
  const float *G1, *g1b, *G2, *g2b;   // nn.layers.0, nn.layers.1 with batch_norm folded
  const float* sc;                    // [1 + eps]
  const float *tG1, *tG2;             // tcgen05 [hi|lo] images
};
struct ModelW {
  EncW enc;
  const float *tenc_W1, *tenc_M2, *tenc_C2;            // tcgen05 [hi|lo] images of the encoder matrices
  const float *henc_W1, *henc_M2, *henc_C2, *henc_sc;  // fp16-split images of the same + inverse weight scales (tc_mlp16.cu)
  const float *hpg_P1h, *hpg_P1e, *hpg_P2, *hpg_sc;    // ... of the global pair MLP (P1e: unscaled lo)
  const float *hpl_P1h, *hpl_P1e, *hpl_P2, *hpl_sc;    // ... of the local pair MLP
  const float *tpg_P1h, *tpg_P1e, *tpg_P2;              // ... of the global pair MLP
  const float *tpl_P1h, *tpl_P1e, *tpl_P2;              // ... of the local pair MLP
  const float* sch_emb;   // [100][128], max_norm renorm pre-applied
  BlkW blk[MAX_BLOCKS];
  PairW pg, pl;
  const float* gin_emb;   // [100][128]
  GinW gin[MAX_GIN];
};

// ------------------------------------------------------------------ per-batch device state
struct BatchDev {
  int n_atoms, n_mols, n_static, n_local;
  int64_t cap;
  // topology (caller-owned)
  const int *atom_type, *mol_ptr, *atom_mol;
  const int64_t* mol_gid;
  const int *st_src, *st_dst, *st_type, *st_in_ptr;
  const int *lc_src, *lc_dst, *lc_type, *lc_in_ptr, *lc_canon, *lc_out_ptr, *lc_cdst;
  // edge builder scratch
  unsigned *adj, *adjT;          // [N][mw]
  int mw;                        // 32-bit words per adjacency row: 8 (every molecule <= 256 atoms) or 16 (<= AGD_MAX_MOL_ATOMS)
  int *in_deg, *out_deg;         // [N]
  int *in_ptr, *out_ptr;         // [N+1]
  int* counters;                 // [0]=n_edges [1]=step index [2]=first NaN step
  int* nan_mol;                  // [n_mols] first NaN step of each molecule (agd_nan_steps)
  // radius-extended edges, CSC order (grouped by destination)
  int *e_src, *e_dst, *e_type, *e_canon;
  float* e_len;
  // the same edges in canonical (row-major) order
  int *c_src, *c_dst, *c_type;
  float* c_len;
  float *s_csc, *s_canon;        // edge_inv_global
  // local edges
  float *lc_len, *lcc_len;       // CSC / canonical
  const float* lc_len_in;        // caller-supplied local edge lengths (CSC order) or nullptr: computed from pos
  float *sl_csc, *sl_canon;      // edge_inv_local
  float* ea_loc;                 // [n_local][128]
  // activations
  float* g2;                     // [cap][128]  encoder hidden state feeding filters and the pair MLP
  uint4* g2h;                    // [ceil(cap/128)*128][128 words] fp16-split copy of g2 (AGD_MODE_F16; layout: g2h_index in tc_common.cuh)
  float* cw_all;                 // [2 * num_convs][cap] envelope * distance weight per CFConv layer (AGD_MODE_F16)
  float* filt;                   // [cap][192]  CFConv filters of the current block (conv1 | conv2)
  float *h, *xcat, *agg;         // [N][128], [N][192], [N][192]
  // pair mode of the local branch: local edges come in both directions with identical inputs (length, type, h_src * h_dst), so the
  // edge encoder and the pair MLP run once per undirected pair (representative orientation src < dst; an edge without a twin is its own pair)
  int n_pairs;                   // 0: unavailable
  int pairs_used;                // the last evaluation ran in pair mode (ea_loc holds pair rows)
  int *lp_of;                    // [n_local] pair of each CSC local edge
  int *lp_src, *lp_dst, *lp_type, *lp_ident;   // [n_pairs] representative edges; lp_ident[p] = p
  float *lp_len, *lp_s, *lp_scratch;           // [n_pairs] length / score per pair; write-only twin for the kernels' second output order
  const int* lc_ea_idx;          // GIN: row of ea_loc for each CSC local edge (pair mode) or nullptr
  float *gx0, *gx1;              // [N][128] GIN ping-pong
  float* hmax;                   // [N] per-atom max |h| feeding the pair kernels' per-row scale (AGD_MODE_F16)
  // per-step schedule (device copy)
  float* sched;                  // [n_steps][4] sigma, step_size, noise_scale, use_global
  int sched_cap;
};

struct StepParams {
  float w_global, clip, clip_local, clip_pos;
  uint64_t seed;
  const float* noise;
  float* traj;
  int use_global;   // which per-step graph this launch belongs to
  int step_offset;  // added to the device step counter for the Philox stream
};

// optional per-launch CUDA-event trace (agd_profile_forward)
struct Prof {
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> label;
};

struct LaunchCtx {
  cudaStream_t stream;
  int num_sms;
  int64_t* launch_counter;
  Prof* prof;
  int use_tc;      // 0: fp32 FFMA tile kernels, 1: tcgen05 3xTF32 everywhere, 2: tcgen05 with 3xFP16 two-slot filter kernels
  int f16_debug_filt;   // fused kernels also write the filter tensor
  unsigned long long* f16_timing;   // diagnostics: per-phase cycle counters of the f16 filter kernels (device, 64 values) or nullptr
  int f16_mlp;     // use_tc == 2: edge encoder on the fp16 two-slot kernels (tc_mlp16.cu)
  int f16_pair;    // use_tc == 2: pair MLPs on the fp16 two-slot kernels
  int local_pairs; // local branch: encoder + pair MLP once per undirected pair (default 1)
  int f16_node;    // use_tc == 2: SchNet node chain on the fp16 kernel with double-buffered weight streaming (tc_node16.cu)
  int f16_fuse;    // use_tc == 2: both CFConv layers of a block + the aggregation in one launch (tc_cfconv.cu; no filt tensor, no aggregate kernel)
  int mlp_act;     // activation of the pair MLPs (AGD_ACT_*); anything but relu runs them on the fp32 FFMA kernel
  float cutoff;
  int smooth;
  int num_convs, num_convs_local;
};

// edges.cu
void launch_build_edges(const LaunchCtx& c, const BatchDev& b, const float* pos);
void launch_export_edges(const LaunchCtx& c, const BatchDev& b, const agd_forward_out& out, bool with_scores);
int launch_extend_bond_order(cudaStream_t s, const int* mol_ptr, int n_mols, int n_atoms, const int* bond_ptr,
                             const int* bond_dst, const int* bond_type, int order, int num_bond_types, int* out_count,
                             const int* out_ptr, int* out_dst, int* out_type, int mw);
// encoder.cu
void launch_encoder_global(const LaunchCtx& c, const BatchDev& b, const ModelW& w);
void launch_encoder_local(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos);
void launch_pair_global(const LaunchCtx& c, const BatchDev& b, const ModelW& w);
void launch_pair_local(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* h_local);
// schnet.cu
void launch_filters(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk);
void launch_filters_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk);   // tc_filter.cu
void launch_edge_weights_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w);   // tc_filter16.cu, once per evaluation
void launch_cfconv_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk);   // tc_cfconv.cu: conv1 + conv2 + aggregation
void launch_filters_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk);  // tc_filter16.cu
int f16_lo_shift();   // S of the fp16 lo' = (x - hi) * 2^S split (0: unscaled), pack.py must build the weight images with it
int f16_fuse_default();    // AGD_F16_FUSE (default 1): launch_filters_f16 also performs the CFConv aggregation into agg
bool filters_tc_fused();   // true: launch_filters_tc also performs the CFConv aggregation into agg
// tc_mlp.cu
void launch_encoder_global_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w);
void launch_encoder_local_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos);
void launch_pair_global_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w);
void launch_pair_local_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* h_local);
void launch_encoder_global_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w);   // tc_mlp16.cu
void launch_encoder_local_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos);
void launch_pair_global_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w);
void launch_pair_local_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* h_local);
void launch_schnet_node_f16(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk);   // tc_node16.cu
void launch_schnet_node_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk);   // tc_node.cu
void launch_gin_layer_tc(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int layer, const float* x_in, float* x_out);
void launch_aggregate(const LaunchCtx& c, const float* x, const float* W, const int* src, const int* in_ptr, int n_nodes,
                      int F, float* out);
void launch_schnet_node(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int blk /* -1: embedding + first lin1 */);
// gin.cu
void launch_gin_embed(const LaunchCtx& c, const BatchDev& b, const ModelW& w);
void launch_gin_layer(const LaunchCtx& c, const BatchDev& b, const ModelW& w, int layer, const float* x_in, float* x_out);
// step.cu
void launch_step(const LaunchCtx& c, const BatchDev& b, float* pos, const StepParams& p);
void launch_advance(const LaunchCtx& c, const BatchDev& b);
void launch_gin_message(cudaStream_t s, const float* x, const float* ea, const int* ea_idx, const int* src, const int* in_ptr, int n_nodes,
                        float eps, const float* one_plus_eps_dev, float* out);
void launch_local_pairs_expand(const LaunchCtx& c, const BatchDev& b);
void launch_kabsch_rmsd(cudaStream_t s, const float* ref, const float* gen, const int* sel, int n_sel, int n_atoms, int n_ref, int n_gen,
                        float* out);
void launch_gather_rows128(cudaStream_t s, const float* src, const int* idx, int n, float* dst);
void launch_eq_transform_segments(cudaStream_t s, const float* pos, const float* s_out, const int* col_of_out, const int* out_ptr,
                                  const float* s_in, const int* row_of_in, const int* in_ptr, int n_nodes, float* out);
void launch_eq_transform(cudaStream_t s, const float* score, const float* pos, const int* src, const int* dst,
                         const float* len, int64_t n_edges, int n_nodes, float* out);

// bookkeeping after every kernel launch: count it and, when tracing, drop an event behind it
inline void note_launch(const LaunchCtx& c, const char* label) {
  *c.launch_counter += 1;
  if (c.prof) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c.stream);
    c.prof->ev.push_back(e);
    c.prof->label.push_back(label);
  }
}

}  // namespace agd
