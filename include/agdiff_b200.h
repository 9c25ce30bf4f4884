/* agdiff_b200 -- C ABI of the B200-native AGDIFF sampling hot path.
 *
 * The reference (ADicksonLab/AGDIFF) has NO plugin/FFI layer: its boundary for this path is the
 * Python class DualEncoderEpsNetwork (src/agdiff/models/epsnet/dualenc.py:54).  These entry points
 * are what a maintainer would bind from that class (ctypes stub shown in INTEGRATION.md); each one
 * cites the reference code it replaces.  Conventions: plain C, int return codes (0 = ok, <0 =
 * error, text via agd_last_error()), no C++ exceptions cross the boundary, every pointer marked
 * "dev" is a CUDA device pointer on the handle's device, "host" is host memory, all launches go
 * to the caller-supplied cudaStream_t (passed as void*), one handle per device, a handle is not
 * thread-safe.  There is no CPU fallback: every call fails with AGD_ERR_CUDA without a GPU.
 */
#ifndef AGDIFF_B200_H
#define AGDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGD_ABI_VERSION 1
#define AGD_HIDDEN 128            /* hidden_dim; pinned by Linear(256, hidden) at schnet.py:190-192 */
#define AGD_MAX_MOL_ATOMS 512     /* per-molecule limit of the adjacency bit-matrix edge builder (rows of 8 x 32 bits while every
                                   * molecule of the batch has <= 256 atoms, 16 x 32 otherwise)      */
#define AGD_MAX_RADIUS_NBRS 32    /* torch_cluster default used by common.py:217                    */

enum {
  AGD_OK = 0,
  AGD_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  AGD_ERR_CUDA = -2,      /* CUDA runtime error (text in agd_last_error) */
  AGD_ERR_CAPACITY = -3,  /* a caller-supplied buffer or the workspace is too small */
  AGD_ERR_NAN = -4,       /* NaN positions during sampling (FloatingPointError, dualenc.py:539-541) */
  AGD_ERR_RANGE = -5      /* an activation left the fp16-split range in AGD_MODE_F16: nothing valid was produced,
                             re-run the call in AGD_MODE_TF32 (the Python mirror does this automatically) */
};

typedef struct agd_handle agd_handle;   /* model: configuration + packed device weights */
typedef struct agd_batch agd_batch;     /* one batch of molecules: topology + workspace  */

/* configs/{qm9,drugs}_default.yml "model:" block, fields read at dualenc.py:64-98,173-174 */
typedef struct {
  int32_t hidden_dim;        /* must be 128 */
  int32_t num_convs;         /* SchNet interaction blocks (6) */
  int32_t num_convs_local;   /* GIN layers (4) */
  int32_t smooth_conv;       /* 0: gaussian cutoff, 1: cosine cutoff (schnet.py:140-145) */
  float cutoff;              /* 10.0 */
  int32_t device;            /* CUDA device ordinal */
} agd_config;

/* Topology of a batch, int32, device pointers; built once per sampler call by the host mirror.
 * "static" edges = the caller's bond graph after bond-order extension (common.py:135-205 /
 * transforms.py:44-71), coalesced (sorted by row*N+col, duplicate types summed).  They are given
 * twice: in CSC order (grouped by destination = edge_index[1], sources ascending) and by their
 * position in canonical order.  "local" edges are the static edges with type > 0
 * (dualenc.py:566-567); when every static type is > 0 both lists coincide. */
typedef struct {
  int32_t n_atoms, n_mols;
  const int32_t* atom_type;   /* dev [n_atoms]   atomic numbers, < 100 */
  const int32_t* mol_ptr;     /* dev [n_mols+1]  atoms of molecule m = mol_ptr[m] .. mol_ptr[m+1) */
  const int32_t* atom_mol;    /* dev [n_atoms]   molecule of each atom (the sorted `batch` vector) */
  const int64_t* mol_gid;     /* dev [n_mols]    global molecule id (keys the Philox noise stream) */
  /* static edges, CSC order */
  int32_t n_static;
  const int32_t* st_src;      /* dev [n_static] edge_index[0] */
  const int32_t* st_dst;      /* dev [n_static] edge_index[1] */
  const int32_t* st_type;     /* dev [n_static] */
  const int32_t* st_in_ptr;   /* dev [n_atoms+1] */
  /* local edges (type > 0), CSC order + canonical bookkeeping */
  int32_t n_local;
  const int32_t* lc_src;      /* dev [n_local] */
  const int32_t* lc_dst;      /* dev [n_local] */
  const int32_t* lc_type;     /* dev [n_local] */
  const int32_t* lc_in_ptr;   /* dev [n_atoms+1] */
  const int32_t* lc_canon;    /* dev [n_local] position of CSC edge e in canonical (row-major) order */
  const int32_t* lc_out_ptr;  /* dev [n_atoms+1] canonical segments by source atom */
  const int32_t* lc_cdst;     /* dev [n_local] destination atom of canonical local edge */
  int64_t edge_capacity;      /* upper bound on edges after radius extension for this batch */
} agd_batch_desc;

/* langevin_dynamics_sample_diffusion arguments (dualenc.py:441-461).  Per-step scalars are
 * computed by the host mirror in fp32 exactly as dualenc.py:468,515,532-533 do. */
typedef struct {
  int32_t n_steps;
  const float* sigma;        /* host [n_steps] sigmas[i], i = T-1 ... T-n_steps            */
  const float* step_size;    /* host [n_steps] step_lr * (sigma/0.01)^2                    */
  const float* noise_scale;  /* host [n_steps] sqrt(2*step_size)                           */
  const uint8_t* use_global; /* host [n_steps] sigma < global_start_sigma (dualenc.py:515) */
  float w_global;
  float clip;                /* global clip_norm limit (dualenc.py:523) */
  float clip_local;          /* < 0: no local clipping (clip_local=None) */
  float clip_pos;            /* < 0: no clamp (clip_pos=None) */
  uint64_t seed;             /* Philox seed when noise == NULL */
  const float* noise;        /* dev [n_steps][n_atoms][3] injected noise, or NULL */
  float* traj;               /* dev [n_steps][n_atoms][3] positions after each step, or NULL */
  int32_t use_cuda_graph;    /* 1: replay captured per-step graphs, 0: plain launches */
  int32_t step_offset;       /* index of this call's first step within the whole trajectory (Philox counter) */
} agd_sample_params;

/* forward outputs (dualenc.py:241-251), caller-allocated device buffers sized edge_capacity /
 * n_local; canonical (row-major sorted) edge order, int32 indices (the host mirror widens to
 * int64).  Any pointer may be NULL to skip that output. */
typedef struct {
  float* edge_inv_global;   /* dev [cap]      */
  float* edge_inv_local;    /* dev [n_local]  canonical order of the local edges */
  int32_t* edge_row;        /* dev [cap] edge_index[0] */
  int32_t* edge_col;        /* dev [cap] edge_index[1] */
  int32_t* edge_type;       /* dev [cap] */
  float* edge_length;       /* dev [cap] */
  int32_t* n_edges;         /* dev [1]   */
} agd_forward_out;

int agd_abi_version(void);
const char* agd_last_error(void);

/* get_model(config) (epsnet/__init__.py:4-8) + .to(device) */
int agd_create(const agd_config* cfg, agd_handle** out);
void agd_destroy(agd_handle* h);

/* The packed-weight layout is owned by the library: slot i has a name and a float count; the host
 * mirror folds BatchNorm(eval), merges back-to-back Linears and transposes in fp64, then uploads.
 * Replaces load_state_dict + eval() for this path (scripts/test.py:111-114). */
int agd_weight_slot_count(const agd_handle* h);
const char* agd_weight_slot_name(const agd_handle* h, int slot);
int64_t agd_weight_slot_size(const agd_handle* h, int slot);
int agd_load_weights(agd_handle* h, const float* packed_host, const int64_t* slot_offsets, int n_slots,
                     int64_t n_floats);

/* workspace is owned by the library (cudaMalloc); size query for planning */
int64_t agd_batch_workspace_bytes(const agd_handle* h, const agd_batch_desc* d);
int agd_batch_create(agd_handle* h, const agd_batch_desc* d, agd_batch** out);
void agd_batch_destroy(agd_batch* b);

/* extend_graph_order_radius + get_distance (common.py:236-264, geometry.py:5-6): builds the
 * canonical edge list for `pos`; outputs as in agd_forward_out (edge_inv_* ignored). */
int agd_build_edges(agd_handle* h, agd_batch* b, const float* pos_dev, const agd_forward_out* out, void* stream);

/* DualEncoderEpsNetwork.forward(..., return_edges=True, extend_radius=True) (dualenc.py:142-251)
 * for a batch whose static edges are already order-extended. */
int agd_forward(agd_handle* h, agd_batch* b, const float* pos_dev, const agd_forward_out* out, void* stream);

/* forward with caller-supplied edges (dualenc.py:166: edges are only rebuilt when one of edge_index / edge_type /
 * edge_length is None) and for extend_radius=False.  All arrays are device pointers; the edge list is given in CSC order
 * (grouped by edge_index[1], sources ascending) with its canonical bookkeeping, exactly what agd_build_edges produces
 * internally; the host mirror derives it from the caller's canonical list.  The batch's local lists (lc_*) must be the
 * type > 0 subset of this edge list. */
typedef struct {
  int32_t n_edges;
  const int32_t *e_src, *e_dst, *e_type, *e_canon;   /* dev [n_edges], CSC order; e_canon = position in the caller's order */
  const float* e_len;                                 /* dev [n_edges] */
  const int32_t *in_ptr, *out_ptr;                    /* dev [n_atoms+1] */
  const int32_t *c_src, *c_dst, *c_type;              /* dev [n_edges], caller's (canonical) order */
  const float* c_len;
  const float* lc_len;                                /* dev [n_local] lengths of the local edges, CSC order */
} agd_edge_set;
int agd_forward_edges(agd_handle* h, agd_batch* b, const float* pos_dev, const agd_edge_set* edges, const agd_forward_out* out,
                      void* stream);

/* langevin_dynamics_sample_diffusion loop body x n_steps (dualenc.py:476-545), in place on pos.
 * first_nan_step (host, may be NULL) receives the first step whose update produced NaN or -1;
 * the call returns AGD_ERR_NAN in that case (the host mirror raises FloatingPointError). */
int agd_sample(agd_handle* h, agd_batch* b, float* pos_dev_inout, const agd_sample_params* p,
               int32_t* first_nan_step, void* stream);
/* After agd_sample: for every molecule of the batch the first step at which its positions became NaN, or -1 (host array of
 * n_mols entries).  Lets a caller that batches many molecules per call repeat only the offenders (the reference's retry loop,
 * scripts/test.py:144-181, is per molecule).  agd_sample itself stops launching steps soon after the first NaN (it polls the
 * device flag every 32 launches) and returns AGD_ERR_NAN. */
int agd_nan_steps(agd_batch* b, int32_t* host_out, int32_t n_mols);

/* _extend_graph_order / AddHigherOrderEdges (common.py:135-205) on device for one batch:
 * CSR bond graph in, per-atom sorted (dst, type) lists out; two passes (count, fill). */
int agd_extend_bond_order(const int32_t* mol_ptr, int32_t n_mols, int32_t n_atoms, const int32_t* bond_ptr,
                          const int32_t* bond_dst, const int32_t* bond_type, int32_t order, int32_t num_bond_types,
                          int32_t* out_count /* dev [n_atoms] */, const int32_t* out_ptr /* dev [n_atoms+1] or NULL */,
                          int32_t* out_dst, int32_t* out_type, void* stream);

/* stand-alone kernels exported for the roofline microbenchmarks (BASELINE.json config 4) and for
 * module-level parity tests; all operate on CSC-sorted edges. */
int agd_op_cfconv_aggregate(const float* x /*dev [N][F]*/, const float* W /*dev [E][F]*/, const int32_t* src,
                            const int32_t* in_ptr, int32_t n_nodes, int32_t F, float* out /*dev [N][F]*/, void* stream);
int agd_op_eq_transform(const float* score /*dev [E]*/, const float* pos, const int32_t* src, const int32_t* dst,
                        const float* length, int64_t n_edges, int32_t n_nodes, float* out /*dev [N][3]*/, void* stream);

/* GIN aggregation, gin.py:76-96 (GINEConv.forward + message): out_i = (1 + eps) * x_i + sum_{e in in(i)} relu(x[src_e] + ea_e)
 * over CSC-sorted edges, 128 features - the gather phase of the fused GIN layer kernel as a kernel of its own. */
int agd_op_gin_message(const float* x /*dev [N][128]*/, const float* ea /*dev [E][128]*/, const int32_t* src /*dev [E]*/,
                       const int32_t* in_ptr /*dev [N+1]*/, int32_t n_nodes, float eps, float* out /*dev [N][128]*/, void* stream);
/* eq_transform, geometry.py:9-17, in the atomics-free form the step kernel uses: the edge list is given twice, sorted by row
 * (out-segments: out_ptr, the col of each edge, its score) and sorted by col (in-segments: in_ptr, the row of each edge, its
 * score); lengths are recomputed from pos. */
int agd_op_eq_transform_segments(const float* pos /*dev [N][3]*/, const float* score_out, const int32_t* col_of_out,
                                 const int32_t* out_ptr, const float* score_in, const int32_t* row_of_in, const int32_t* in_ptr,
                                 int32_t n_nodes, float* out /*dev [N][3]*/, void* stream);

/* COV/MAT building block, utils/evaluation/covmat.py:16-34: out[i][j] = RMSD of generated conformer j onto reference conformer i
 * after optimal superposition (proper rotation) over the selected atoms (sel = NULL: all; the reference compares heavy atoms).
 * Atom order is taken as given - RDKit's GetBestRMS additionally minimises over the molecule's symmetry permutations. */
int agd_op_kabsch_rmsd(const float* ref /*dev [n_ref][n_atoms][3]*/, const float* gen /*dev [n_gen][n_atoms][3]*/,
                       const int32_t* sel /*dev [n_sel] or NULL*/, int32_t n_sel, int32_t n_atoms, int32_t n_ref, int32_t n_gen,
                       float* out /*dev [n_ref][n_gen]*/, void* stream);

/* host-only helper (no CUDA call): the pair map the local branch uses for a CSC-sorted local edge list given as HOST
 * arrays - pair_of[e] in [0, *n_pairs); two edges share a pair iff they are each other's reverse with the same type. */
int agd_host_local_pairs(const int32_t* src, const int32_t* dst, const int32_t* type, const int32_t* in_ptr /*host [n_atoms+1]*/,
                         int32_t n_local, int32_t n_atoms, int32_t* pair_of /*host [n_local]*/, int32_t* n_pairs);

/* debugging / tests: copy an internal per-batch tensor to a caller device buffer.  Names:
 * "g2", "h_global", "h_local", "ea_local", "xcat", "agg", "filt".  Returns element count or <0. */
int64_t agd_debug_fetch(agd_batch* b, const char* name, float* dst_dev, int64_t capacity);

/* measurement: run the per-step network evaluation once (no position update) with a CUDA event
 * behind every kernel launch on the launching stream; ms[i] is the device time of launch i and
 * labels receives the '\n'-separated kernel labels.  n_edges_out (host) gets the edge count. */
int agd_profile_forward(agd_handle* h, agd_batch* b, const float* pos_dev, int32_t with_global, char* labels,
                        int64_t labels_cap, float* ms, int32_t cap, int32_t* n_out, int32_t* n_edges_out, void* stream);

/* number of kernel launches issued by the library since the handle was created */
int64_t agd_launch_count(const agd_handle* h);

/* arithmetic of the dense per-edge / per-atom contractions (all fp32-faithful, all on the GPU):
 *   AGD_MODE_FFMA  fp32 FFMA tile kernels,
 *   AGD_MODE_TF32  tcgen05 kind::tf32 with the 3xTF32 split (unbounded range),
 *   AGD_MODE_F16   tcgen05 kind::f16 with the fp16 hi/lo' split in the CFConv filter kernels, two edge tiles in
 *                  flight per SM (default; activations beyond +-65000 make the call fail with AGD_ERR_RANGE).
 * The environment variable AGD_TC_FILTERS=0/1/2 selects the initial mode of new handles. */
enum { AGD_MODE_FFMA = 0, AGD_MODE_TF32 = 1, AGD_MODE_F16 = 2 };
int agd_set_mode(agd_handle* h, int mode);
int agd_get_mode(const agd_handle* h);
/* tuning / A-B switches (AGD_MODE_F16 only; all default to 1 except the diagnostics):
 *   "f16_fuse"  (env AGD_F16_FUSE)  both CFConv layers of a block (conv1 F=128, conv2 F=64) and the aggregation run in ONE
 *               warp-specialised launch (tc_cfconv.cu: epilogue / aggregation / loader / MMA-issue warps, the encoder state is read
 *               once per block; same sums in the same order as the stand-alone aggregate kernel, bit for bit);
 *               0: one filter kernel per conv, filter tensor to HBM + aggregate kernel
 *   "f16_mlp"   (env AGD_F16_MLP)   edge encoder on the fp16 two-slot kernels (0: 3xTF32 kernels)
 *   "f16_pair"  (env AGD_F16_PAIR)  pair MLPs on the fp16 two-slot kernels
 *   "f16_node"  (env AGD_F16_NODE)  SchNet node chain on the fp16 kernel with double-buffered weight streaming
 *   "mlp_act"   activation of the two pair MLPs (config field mlp_act; common.py:59-62 takes any torch.nn.functional name):
 *               0 relu (default), 1 gelu, 2 silu, 3 tanh, 4 sigmoid, 5 leaky_relu, 6 elu, 7 softplus.  Anything but relu runs the
 *               pair MLPs on the fp32 FFMA kernel and the encoders on the 3xTF32 kernels (the fp16-split pair kernel relies on
 *               relu being positively homogeneous)
 *   "f16_debug_filt" (0)  the fused CFConv kernels also write the filter tensor (tests)
 *   "f16_timing"     (0)  clock64 phase counters of the CFConv kernels (needs a build with -DAGD_F16_TIMING), agd_debug_timing */
int agd_set_option(agd_handle* h, const char* name, int value);
/* 1 if a kernel of the last forward on this batch saw an activation outside the fp16-split range (host sync) */
int agd_range_flag(agd_batch* b, int32_t* flag_out);
/* S of the fp16 split lo' = (x - hi) * 2^S the library was built/configured with; the packer builds the images with it */
int agd_f16_lo_shift(void);
/* diagnostics: after agd_set_option(h, "f16_timing", 1), the accumulated clock64 cycles per pipeline phase of the fp16
 * filter kernels: out64[(group*2 + observer)*8 + phase], summed over CTAs and launches */
int agd_debug_timing(agd_handle* h, uint64_t* out64);

#ifdef __cplusplus
}
#endif
#endif /* AGDIFF_B200_H */
