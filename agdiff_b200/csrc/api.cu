// C ABI of libagdiff_b200.so (see include/agdiff_b200.h): handle/weights/batch management, the
// forward pass, and the sampling loop replayed from two captured CUDA graphs (local-only step and
// local+global step).  No CPU fallback: every entry point needs a CUDA device.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "kernels.h"

namespace agd {
void set_encoder_attributes();
void set_schnet_attributes();
void set_gin_attributes();
void set_tc_attributes();
void set_tc_mlp_attributes();
void set_tc_node_attributes();
void set_tc16_attributes();
void set_tc_mlp16_attributes();
void set_tc_node16_attributes();
void set_tc_cfconv_attributes();
}  // namespace agd

using namespace agd;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(AGD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)

struct Slot {
  std::string name;
  int64_t size;
  const float** field;
};

struct agd_handle {
  agd_config cfg;
  ModelW w;
  std::vector<Slot> slots;
  float* dev_weights = nullptr;
  int64_t n_weights = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_poll[2] = {nullptr, nullptr};
  int* poll_flag = nullptr;   // pinned host copies [2] of the first-NaN-step counter (early exit of agd_sample)
  int64_t launches = 0;
  int f16_fuse = 1;
  int f16_mlp = 1;
  int f16_pair = 1;
  int f16_node = 1;
  int local_pairs = 1;
  int f16_debug_filt = 0;
  unsigned long long* f16_timing = nullptr;
  int mlp_act = 0;  // AGD_ACT_* of the pair MLPs (agd_set_option "mlp_act")
  int use_tc = 2;   // AGD_TC_FILTERS / agd_set_mode: 0 FFMA, 1 tcgen05 3xTF32, 2 tcgen05 + 3xFP16 filter kernels
};

struct agd_batch {
  agd_handle* h;
  BatchDev d;
  void* slab = nullptr;
  int64_t slab_bytes = 0;
};

static void build_slots(agd_handle* h) {
  auto& s = h->slots;
  ModelW& w = h->w;
  auto add = [&](const std::string& n, int64_t sz, const float** f) { s.push_back({n, sz, f}); };
  const int H = HID;
  add("enc.fe_w", H, &w.enc.fe_w);
  add("enc.fe_b", H, &w.enc.fe_b);
  add("enc.T1", 100 * H, &w.enc.T1);
  add("enc.W1", H * H, &w.enc.W1);
  add("enc.T2", 100 * H, &w.enc.T2);
  add("enc.M2", H * H, &w.enc.M2);
  add("enc.C2", H * H, &w.enc.C2);
  add("enc.c2b", H, &w.enc.c2b);
  add("tenc.W1", 2 * H * H, &w.tenc_W1); add("tenc.M2", 2 * H * H, &w.tenc_M2); add("tenc.C2", 2 * H * H, &w.tenc_C2);
  add("tpg.P1h", 2 * H * H, &w.tpg_P1h); add("tpg.P1e", 2 * H * H, &w.tpg_P1e); add("tpg.P2", 2 * 64 * H, &w.tpg_P2);
  add("tpl.P1h", 2 * H * H, &w.tpl_P1h); add("tpl.P1e", 2 * H * H, &w.tpl_P1e); add("tpl.P2", 2 * 64 * H, &w.tpl_P2);
  add("henc.W1", H * H, &w.henc_W1); add("henc.M2", H * H, &w.henc_M2); add("henc.C2", H * H, &w.henc_C2); add("henc.sc", 4, &w.henc_sc);
  add("hpg.P1h", H * H, &w.hpg_P1h); add("hpg.P1e", H * H, &w.hpg_P1e); add("hpg.P2", 64 * H, &w.hpg_P2); add("hpg.sc", 4, &w.hpg_sc);
  add("hpl.P1h", H * H, &w.hpl_P1h); add("hpl.P1e", H * H, &w.hpl_P1e); add("hpl.P2", 64 * H, &w.hpl_P2); add("hpl.sc", 4, &w.hpl_sc);
  add("sch.emb", 100 * H, &w.sch_emb);
  for (int k = 0; k < h->cfg.num_convs; ++k) {
    BlkW& b = w.blk[k];
    const std::string p = "blk" + std::to_string(k) + ".";
    add(p + "L1a", H * 128, &b.L1a);   add(p + "l1ab", 128, &b.l1ab);
    add(p + "L1b", H * 64, &b.L1b);    add(p + "l1bb", 64, &b.l1bb);
    add(p + "F1a", H * 128, &b.F1a);   add(p + "f1ab", 128, &b.f1ab);
    add(p + "F2a", 128 * 128, &b.F2a); add(p + "f2ab", 128, &b.f2ab);
    add(p + "F1b", H * 64, &b.F1b);    add(p + "f1bb", 64, &b.f1bb);
    add(p + "F2b", 64 * 64, &b.F2b);   add(p + "f2bb", 64, &b.f2bb);
    add(p + "dw1", 128, &b.dw1);       add(p + "dw2", 128, &b.dw2);
    add(p + "L2a", 128 * H, &b.L2a);   add(p + "l2ab", H, &b.l2ab);
    add(p + "L2b", 64 * H, &b.L2b);    add(p + "l2bb", H, &b.l2bb);
    add(p + "LIN", 256 * H, &b.LIN);   add(p + "linb", H, &b.linb);
    add(p + "A1", H * 64, &b.A1);      add(p + "a1b", 64, &b.a1b);
    add(p + "a2w", 64, &b.a2w);
    add(p + "S1", H * 8, &b.S1);       add(p + "S2", 8 * H, &b.S2);
    add(p + "sc", 4, &b.sc);
    add(p + "tF1a", 2 * 128 * 128, &b.tF1a); add(p + "tF2a", 2 * 128 * 128, &b.tF2a);
    add(p + "tF1b", 2 * 64 * 128, &b.tF1b);   add(p + "tF2b", 2 * 64 * 64, &b.tF2b);
    add(p + "tL1a", 2 * 128 * 128, &b.tL1a); add(p + "tL1b", 2 * 64 * 128, &b.tL1b);
    add(p + "tL2a", 2 * 128 * 128, &b.tL2a); add(p + "tL2b", 2 * 128 * 64, &b.tL2b);
    add(p + "tLINa", 2 * 128 * 128, &b.tLINa); add(p + "tLINb", 2 * 128 * 128, &b.tLINb);
    add(p + "tA1", 2 * 64 * 128, &b.tA1);
    add(p + "hF1a", 128 * 128, &b.hF1a); add(p + "hF2a", 128 * 128, &b.hF2a);
    add(p + "hF1b", 64 * 128, &b.hF1b);   add(p + "hF2b", 64 * 64, &b.hF2b);
    add(p + "hsc", 4, &b.hsc);
    add(p + "hL2a", 128 * 128, &b.hL2a); add(p + "hLINa", 128 * 128, &b.hLINa); add(p + "hL2b", 128 * 64, &b.hL2b);
    add(p + "hLINb", 128 * 128, &b.hLINb); add(p + "hA1", 64 * 128, &b.hA1);
    add(p + "hL1a", 128 * 128, &b.hL1a); add(p + "hL1b", 64 * 128, &b.hL1b); add(p + "hnsc", 8, &b.hnsc);
  }
  auto add_pair = [&](const std::string& p, PairW& q) {
    add(p + "P1h", H * H, &q.P1h); add(p + "P1e", H * H, &q.P1e); add(p + "p1b", H, &q.p1b);
    add(p + "P2", H * 64, &q.P2);  add(p + "p2b", 64, &q.p2b);
    add(p + "p3w", 64, &q.p3w);    add(p + "p3b", 1, &q.p3b);
  };
  add_pair("pg.", w.pg);
  add_pair("pl.", w.pl);
  add("gin.emb", 100 * H, &w.gin_emb);
  for (int k = 0; k < h->cfg.num_convs_local; ++k) {
    GinW& g = w.gin[k];
    const std::string p = "gin" + std::to_string(k) + ".";
    add(p + "G1", H * H, &g.G1); add(p + "g1b", H, &g.g1b);
    add(p + "G2", H * H, &g.G2); add(p + "g2b", H, &g.g2b);
    add(p + "sc", 1, &g.sc);
    add(p + "tG1", 2 * H * H, &g.tG1); add(p + "tG2", 2 * H * H, &g.tG2);
  }
}

static LaunchCtx make_ctx(agd_handle* h) {
  LaunchCtx c;
  c.stream = h->stream;
  c.num_sms = h->num_sms;
  c.launch_counter = &h->launches;
  c.prof = nullptr;
  // mlp_act != relu: the pair MLPs run on the fp32 FFMA kernel, which reads the fp32 encoder state - written by the 3xTF32 /
  // FFMA encoders only - so the fp16-split mode is not used for such a model
  c.mlp_act = h->mlp_act;
  c.use_tc = (h->mlp_act != 0 && h->use_tc == 2) ? 1 : h->use_tc;
  c.f16_fuse = h->f16_fuse;
  c.f16_mlp = (c.use_tc == 2) ? h->f16_mlp : 0;
  c.f16_pair = (c.use_tc == 2) ? h->f16_pair : 0;
  c.f16_node = (c.use_tc == 2) ? h->f16_node : 0;
  c.local_pairs = h->local_pairs;
  c.f16_debug_filt = h->f16_debug_filt;
  c.f16_timing = h->f16_timing;
  c.cutoff = h->cfg.cutoff;
  c.smooth = h->cfg.smooth_conv;
  c.num_convs = h->cfg.num_convs;
  c.num_convs_local = h->cfg.num_convs_local;
  return c;
}

// order library work after everything already queued on the caller's stream, and back
static int enter(agd_handle* h, void* user_stream) {
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaEventRecord(h->ev_in, (cudaStream_t)user_stream));
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_in, 0));
  return AGD_OK;
}
static int leave(agd_handle* h, void* user_stream) {
  CUDA_TRY(cudaEventRecord(h->ev_out, h->stream));
  CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)user_stream, h->ev_out, 0));
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

// NVTX range per phase of a network evaluation (host side: visible in Nsight Systems timelines around the launches; inside a
// captured graph they bracket the capture, not the replays)
struct Nvtx {
  explicit Nvtx(const char* name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
};

// ------------------------------------------------------------------ launch sequences
// Local branch (dualenc.py:212-236).  Pair mode: the local edge list holds every bond / 2-hop / 3-hop pair in both directions, and
// everything the edge encoder and the pair MLP see is symmetric in the direction - |pos_i - pos_j| (squares of negated
// differences: the same bits), the edge type, h_i * h_j - so both run once per undirected pair over a "view" of the batch whose
// local edge list is the list of representative edges, and their results are handed to both directions afterwards
// (bit-identical to evaluating every directed edge).  The GIN gather in between reads edge_attr through the pair map.
static void run_local_branch(const LaunchCtx& c, const BatchDev& b_in, const ModelW& w, const float* pos, const float** h_out) {
  Nvtx r_all("agd.local_branch");
  const bool pairs = c.local_pairs && b_in.n_pairs > 0 && b_in.lc_len_in == nullptr;
  const_cast<BatchDev&>(b_in).pairs_used = pairs ? 1 : 0;
  BatchDev v = b_in;   // what the encoder and the pair MLP see
  BatchDev b = b_in;   // what the GIN layers see
  if (pairs) {
    v.n_local = b_in.n_pairs;
    v.lc_src = b_in.lp_src; v.lc_dst = b_in.lp_dst; v.lc_type = b_in.lp_type; v.lc_canon = b_in.lp_ident;
    v.lc_len = b_in.lp_len; v.lcc_len = b_in.lp_scratch; v.sl_csc = b_in.lp_s; v.sl_canon = b_in.lp_scratch;
    b.lc_ea_idx = b_in.lp_of;
  }
  {
  Nvtx r("agd.edge_encoder.local");
  if (c.f16_mlp) launch_encoder_local_f16(c, v, w, pos);
  else if (c.use_tc) launch_encoder_local_tc(c, v, w, pos);
  else launch_encoder_local(c, v, w, pos);
  }
  const float* x_in = b.gx0;
  float* x_out = b.gx1;
  {
    Nvtx r("agd.gin");
    launch_gin_embed(c, b, w);
    for (int k = 0; k < c.num_convs_local; ++k) {
      if (c.use_tc) launch_gin_layer_tc(c, b, w, k, x_in, x_out); else launch_gin_layer(c, b, w, k, x_in, x_out);
      const float* t = x_in;
      x_in = x_out;
      x_out = const_cast<float*>(t);
    }
  }
  Nvtx r_pair("agd.pair_mlp.local");
  if (c.f16_pair) launch_pair_local_f16(c, v, w, x_in);
  else if (c.use_tc && c.mlp_act == 0) launch_pair_local_tc(c, v, w, x_in);
  else launch_pair_local(c, v, w, x_in);
  if (pairs) launch_local_pairs_expand(c, b_in);
  if (h_out) *h_out = x_in;
}

static void run_global_branch(const LaunchCtx& c, const BatchDev& b, const ModelW& w, const float* pos, bool build = true) {
  Nvtx r_all("agd.global_branch");
  if (build) {
    Nvtx r("agd.edges");
    launch_build_edges(c, b, pos);
  }
  {
    Nvtx r("agd.edge_encoder.global");
    if (c.f16_mlp) launch_encoder_global_f16(c, b, w);
    else if (c.use_tc) launch_encoder_global_tc(c, b, w);
    else launch_encoder_global(c, b, w);
  }
  nvtxRangePushA("agd.schnet");
  if (c.use_tc == 2) launch_edge_weights_f16(c, b, w);
  if (c.f16_node) launch_schnet_node_f16(c, b, w, -1);
  else if (c.use_tc) launch_schnet_node_tc(c, b, w, -1);
  else launch_schnet_node(c, b, w, -1);
  for (int k = 0; k < c.num_convs; ++k) {
    Nvtx r_blk("agd.schnet.block");
    launch_filters(c, b, w, k);
    if (!((c.use_tc == 1 && filters_tc_fused()) || (c.use_tc == 2 && c.f16_fuse))) launch_aggregate(c, b.xcat, b.filt, b.e_src, b.in_ptr, b.n_atoms, 192, b.agg);
    if (c.f16_node) launch_schnet_node_f16(c, b, w, k);
    else if (c.use_tc) launch_schnet_node_tc(c, b, w, k);
    else launch_schnet_node(c, b, w, k);
  }
  nvtxRangePop();   // agd.schnet
  Nvtx r_pair("agd.pair_mlp.global");
  if (c.f16_pair) launch_pair_global_f16(c, b, w);
  else if (c.use_tc && c.mlp_act == 0) launch_pair_global_tc(c, b, w);
  else launch_pair_global(c, b, w);
}

extern "C" {

int agd_abi_version(void) { return AGD_ABI_VERSION; }
const char* agd_last_error(void) { return g_err.c_str(); }

int agd_create(const agd_config* cfg, agd_handle** out) {
  if (!cfg || !out) return fail(AGD_ERR_INVALID, "null argument");
  if (cfg->hidden_dim != HID) return fail(AGD_ERR_INVALID, "hidden_dim must be 128 (schnet.py:190-192 pins Linear(256, hidden))");
  if (cfg->num_convs < 1 || cfg->num_convs > MAX_BLOCKS) return fail(AGD_ERR_INVALID, "num_convs out of range");
  if (cfg->num_convs_local < 1 || cfg->num_convs_local > MAX_GIN) return fail(AGD_ERR_INVALID, "num_convs_local out of range");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(AGD_ERR_CUDA, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(cfg->device));
  agd_handle* h = new agd_handle();
  h->cfg = *cfg;
  std::memset(&h->w, 0, sizeof(h->w));
  build_slots(h);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  h->num_sms = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_poll[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&h->ev_poll[1], cudaEventDisableTiming));
  CUDA_TRY(cudaMallocHost(&h->poll_flag, 2 * sizeof(int)));
  set_encoder_attributes();
  set_schnet_attributes();
  set_gin_attributes();
  set_tc_attributes();
  set_tc_mlp_attributes();
  set_tc_node_attributes();
  set_tc16_attributes();
  set_tc_mlp16_attributes();
  set_tc_node16_attributes();
  set_tc_cfconv_attributes();
  h->f16_fuse = f16_fuse_default();
  if (const char* e = std::getenv("AGD_F16_MLP")) h->f16_mlp = (e[0] != '0');
  if (const char* e = std::getenv("AGD_F16_PAIR")) h->f16_pair = (e[0] != '0');
  if (const char* e = std::getenv("AGD_F16_NODE")) h->f16_node = (e[0] != '0');
  if (const char* e = std::getenv("AGD_LOCAL_PAIRS")) h->local_pairs = (e[0] != '0');
  if (const char* e = std::getenv("AGD_F16_DEBUG")) h->f16_debug_filt = std::atoi(e);
  if (const char* e = std::getenv("AGD_TC_FILTERS")) h->use_tc = (e[0] == '0') ? 0 : (e[0] == '1') ? 1 : 2;
  CUDA_TRY(cudaGetLastError());
  *out = h;
  return AGD_OK;
}

void agd_destroy(agd_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->dev_weights) cudaFree(h->dev_weights);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  if (h->ev_out) cudaEventDestroy(h->ev_out);
  if (h->ev_poll[0]) cudaEventDestroy(h->ev_poll[0]);
  if (h->ev_poll[1]) cudaEventDestroy(h->ev_poll[1]);
  if (h->poll_flag) cudaFreeHost(h->poll_flag);
  delete h;
}

int agd_weight_slot_count(const agd_handle* h) { return h ? (int)h->slots.size() : 0; }
const char* agd_weight_slot_name(const agd_handle* h, int slot) {
  if (!h || slot < 0 || slot >= (int)h->slots.size()) return nullptr;
  return h->slots[slot].name.c_str();
}
int64_t agd_weight_slot_size(const agd_handle* h, int slot) {
  if (!h || slot < 0 || slot >= (int)h->slots.size()) return -1;
  return h->slots[slot].size;
}

int agd_load_weights(agd_handle* h, const float* packed_host, const int64_t* slot_offsets, int n_slots, int64_t n_floats) {
  if (!h || !packed_host || !slot_offsets) return fail(AGD_ERR_INVALID, "null argument");
  if (n_slots != (int)h->slots.size()) return fail(AGD_ERR_INVALID, "slot count mismatch");
  for (int i = 0; i < n_slots; ++i) {
    if (slot_offsets[i] < 0 || slot_offsets[i] + h->slots[i].size > n_floats || (slot_offsets[i] % 4) != 0)
      return fail(AGD_ERR_INVALID, "bad offset for slot " + h->slots[i].name);
  }
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->n_weights != n_floats) {
    if (h->dev_weights) CUDA_TRY(cudaFree(h->dev_weights));
    h->dev_weights = nullptr;
    CUDA_TRY(cudaMalloc(&h->dev_weights, sizeof(float) * (size_t)n_floats));
    h->n_weights = n_floats;
  }
  CUDA_TRY(cudaMemcpy(h->dev_weights, packed_host, sizeof(float) * (size_t)n_floats, cudaMemcpyHostToDevice));
  for (int i = 0; i < n_slots; ++i) *h->slots[i].field = h->dev_weights + slot_offsets[i];
  return AGD_OK;
}

}  // extern "C"

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct Carver {
  char* base;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

static void carve(BatchDev& d, Carver& c) {
  const size_t N = (size_t)d.n_atoms, E = (size_t)(d.cap > 0 ? d.cap : 1), L = (size_t)(d.n_local > 0 ? d.n_local : 1);
  d.adj = c.take<unsigned>(N * (size_t)d.mw);
  d.adjT = c.take<unsigned>(N * (size_t)d.mw);
  d.in_deg = c.take<int>(N);
  d.out_deg = c.take<int>(N);
  d.in_ptr = c.take<int>(N + 1);
  d.out_ptr = c.take<int>(N + 1);
  d.counters = c.take<int>(8);
  d.nan_mol = c.take<int>((size_t)(d.n_mols > 0 ? d.n_mols : 1));
  d.e_src = c.take<int>(E);
  d.e_dst = c.take<int>(E);
  d.e_type = c.take<int>(E);
  d.e_canon = c.take<int>(E);
  d.e_len = c.take<float>(E);
  d.c_src = c.take<int>(E);
  d.c_dst = c.take<int>(E);
  d.c_type = c.take<int>(E);
  d.c_len = c.take<float>(E);
  d.s_csc = c.take<float>(E);
  d.s_canon = c.take<float>(E);
  d.lc_len = c.take<float>(L);
  d.lcc_len = c.take<float>(L);
  d.sl_csc = c.take<float>(L);
  d.sl_canon = c.take<float>(L);
  d.ea_loc = c.take<float>(L * HID);
  d.g2 = c.take<float>(E * HID);
  d.g2h = c.take<uint4>(((E + TM - 1) / TM) * TM * (HID / 4));
  d.cw_all = c.take<float>(E * 2 * MAX_BLOCKS);
  d.filt = c.take<float>(E * 192);
  d.h = c.take<float>(N * HID);
  d.xcat = c.take<float>(N * 192);
  d.agg = c.take<float>(N * 192);
  d.gx0 = c.take<float>(N * HID);
  d.gx1 = c.take<float>(N * HID);
  d.lp_of = c.take<int>(L);
  d.lp_src = c.take<int>(L);
  d.lp_dst = c.take<int>(L);
  d.lp_type = c.take<int>(L);
  d.lp_ident = c.take<int>(L);
  d.lp_len = c.take<float>(L);
  d.lp_s = c.take<float>(L);
  d.lp_scratch = c.take<float>(L);
  d.hmax = c.take<float>(N);
}

static int check_desc(const agd_batch_desc* d) {
  if (!d) return fail(AGD_ERR_INVALID, "null batch descriptor");
  if (d->n_atoms <= 0 || d->n_mols <= 0) return fail(AGD_ERR_INVALID, "empty batch");
  if (d->edge_capacity < 0 || d->edge_capacity > (int64_t)INT_MAX - TM) return fail(AGD_ERR_CAPACITY, "edge capacity exceeds int32 indexing; split the batch");
  if (!d->atom_type || !d->mol_ptr || !d->atom_mol || !d->mol_gid || !d->st_in_ptr || !d->lc_in_ptr || !d->lc_out_ptr)
    return fail(AGD_ERR_INVALID, "missing topology pointer");
  return AGD_OK;
}

extern "C" {

// 32-bit words per adjacency row for a batch: 8 while every molecule has <= 256 atoms, 16 up to AGD_MAX_MOL_ATOMS; -1 beyond
// (mol_ptr is a device array; this runs once per batch / per bond-order call, never on the step path).
static int adjacency_words(const int32_t* mol_ptr_dev, int n_mols, int* largest_out = nullptr) {
  std::vector<int32_t> hp((size_t)n_mols + 1);
  if (cudaMemcpy(hp.data(), mol_ptr_dev, sizeof(int32_t) * hp.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  int largest = 0;
  for (int m = 0; m < n_mols; ++m) largest = std::max(largest, hp[m + 1] - hp[m]);
  if (largest_out) *largest_out = largest;
  if (largest > AGD_MAX_MOL_ATOMS) return -1;
  return largest <= 256 ? 8 : 16;
}

// Pair map of the local edges (CSC order: grouped by destination, sources ascending): edge (j -> i) with j > i shares the pair of
// its twin (i -> j) when that exists with the same type; every other edge is the representative of its own pair.  Host arrays in,
// host vectors out (no CUDA: also exported as agd_host_local_pairs for the CPU tests).
static void local_pair_map(const int* src, const int* dst, const int* typ, const int* ptr, size_t L, size_t N, std::vector<int>& of,
                           std::vector<int>& ps, std::vector<int>& pd, std::vector<int>& pt) {
  of.assign(L, -1);
  ps.clear(); pd.clear(); pt.clear();
  ps.reserve(L / 2 + 1); pd.reserve(L / 2 + 1); pt.reserve(L / 2 + 1);
  auto twin = [&](size_t e) -> long {   // the edge dst[e] -> src[e], found in the segment of destination src[e]
    const int j = src[e], i = dst[e];
    if (j < 0 || (size_t)j >= N || ptr[j] < 0 || ptr[j + 1] < ptr[j] || (size_t)ptr[j + 1] > L) return -1;
    const int* lo = src + ptr[j];
    const int* hi = src + ptr[j + 1];
    const int* it = std::lower_bound(lo, hi, i);
    if (it == hi || *it != i) return -1;
    const long t = it - src;
    return (dst[t] == j && typ[t] == typ[e]) ? t : -1;
  };
  auto open_pair = [&](size_t e) {
    of[e] = (int)ps.size();
    ps.push_back(src[e]); pd.push_back(dst[e]); pt.push_back(typ[e]);
  };
  for (size_t e = 0; e < L; ++e)
    if (src[e] < dst[e]) open_pair(e);
  for (size_t e = 0; e < L; ++e) {
    if (of[e] >= 0) continue;
    const long t = (src[e] > dst[e]) ? twin(e) : -1;
    if (t >= 0 && of[t] >= 0 && src[t] < dst[t]) of[e] = of[t];
    else open_pair(e);
  }
}

int agd_host_local_pairs(const int32_t* src, const int32_t* dst, const int32_t* type, const int32_t* in_ptr, int32_t n_local,
                         int32_t n_atoms, int32_t* pair_of, int32_t* n_pairs) {
  if (n_local < 0 || n_atoms < 0 || !n_pairs || (n_local > 0 && (!src || !dst || !type || !in_ptr || !pair_of)))
    return fail(AGD_ERR_INVALID, "bad arguments");
  std::vector<int> of, ps, pd, pt;
  local_pair_map(src, dst, type, in_ptr, (size_t)n_local, (size_t)n_atoms, of, ps, pd, pt);
  for (int32_t e = 0; e < n_local; ++e) pair_of[e] = of[(size_t)e];
  *n_pairs = (int32_t)ps.size();
  return AGD_OK;
}

static int build_local_pairs(BatchDev& v) {
  v.n_pairs = 0;
  const size_t L = (size_t)v.n_local, N = (size_t)v.n_atoms;
  if (L == 0) return AGD_OK;
  std::vector<int> src(L), dst(L), typ(L), ptr(N + 1);
  CUDA_TRY(cudaMemcpy(src.data(), v.lc_src, L * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(dst.data(), v.lc_dst, L * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(typ.data(), v.lc_type, L * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(ptr.data(), v.lc_in_ptr, (N + 1) * 4, cudaMemcpyDeviceToHost));
  std::vector<int> of, ps, pd, pt;
  local_pair_map(src.data(), dst.data(), typ.data(), ptr.data(), L, N, of, ps, pd, pt);
  const size_t P = ps.size();
  std::vector<int> ident(P);
  for (size_t p = 0; p < P; ++p) ident[p] = (int)p;
  CUDA_TRY(cudaMemcpy(v.lp_of, of.data(), L * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.lp_src, ps.data(), P * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.lp_dst, pd.data(), P * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.lp_type, pt.data(), P * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(v.lp_ident, ident.data(), P * 4, cudaMemcpyHostToDevice));
  v.n_pairs = (int)P;
  return AGD_OK;
}

int64_t agd_batch_workspace_bytes(const agd_handle* h, const agd_batch_desc* d) {
  (void)h;
  if (check_desc(d) != AGD_OK) return -1;
  BatchDev t{};
  t.n_atoms = d->n_atoms;
  t.n_mols = d->n_mols;
  t.cap = d->edge_capacity;
  t.n_local = d->n_local;
  t.mw = 16;   // upper bound (the row width is only known once the molecule sizes have been read)
  Carver c{nullptr};
  carve(t, c);
  return (int64_t)c.off;
}

int agd_batch_create(agd_handle* h, const agd_batch_desc* d, agd_batch** out) {
  if (!h || !out) return fail(AGD_ERR_INVALID, "null argument");
  int rc = check_desc(d);
  if (rc != AGD_OK) return rc;
  CUDA_TRY(cudaSetDevice(h->cfg.device));
  agd_batch* b = new agd_batch();
  b->h = h;
  BatchDev& v = b->d;
  std::memset(&v, 0, sizeof(v));
  v.n_atoms = d->n_atoms; v.n_mols = d->n_mols; v.n_static = d->n_static; v.n_local = d->n_local; v.cap = d->edge_capacity;
  v.atom_type = d->atom_type; v.mol_ptr = d->mol_ptr; v.atom_mol = d->atom_mol; v.mol_gid = d->mol_gid;
  v.st_src = d->st_src; v.st_dst = d->st_dst; v.st_type = d->st_type; v.st_in_ptr = d->st_in_ptr;
  v.lc_src = d->lc_src; v.lc_dst = d->lc_dst; v.lc_type = d->lc_type; v.lc_in_ptr = d->lc_in_ptr;
  v.lc_canon = d->lc_canon; v.lc_out_ptr = d->lc_out_ptr; v.lc_cdst = d->lc_cdst;
  {
    int largest = 0;
    v.mw = adjacency_words(d->mol_ptr, d->n_mols, &largest);
    if (v.mw < 0) {
      delete b;
      if (v.mw == -2) return fail(AGD_ERR_CUDA, "cannot read mol_ptr (device pointer expected)");
      return fail(AGD_ERR_CAPACITY, "a molecule of " + std::to_string(largest) + " atoms exceeds AGD_MAX_MOL_ATOMS (" + std::to_string(AGD_MAX_MOL_ATOMS) + ")");
    }
  }
  Carver sz{nullptr};
  carve(v, sz);
  b->slab_bytes = (int64_t)sz.off;
  cudaError_t e = cudaMalloc(&b->slab, sz.off);
  if (e != cudaSuccess) {
    delete b;
    return fail(AGD_ERR_CUDA, std::string("workspace cudaMalloc(") + std::to_string(sz.off) + "): " + cudaGetErrorString(e));
  }
  Carver cv{reinterpret_cast<char*>(b->slab)};
  carve(v, cv);
  e = cudaMemset(v.counters, 0, 8 * sizeof(int));
  if (e != cudaSuccess) {
    cudaFree(b->slab);
    delete b;
    return fail(AGD_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
  }
  if (build_local_pairs(v) != AGD_OK) v.n_pairs = 0;   // (pair mode is an optimisation: without the map the local branch evaluates every directed edge)
  *out = b;
  return AGD_OK;
}

void agd_batch_destroy(agd_batch* b) {
  if (!b) return;
  cudaSetDevice(b->h->cfg.device);
  cudaStreamSynchronize(b->h->stream);
  if (b->d.sched) cudaFree(b->d.sched);
  if (b->slab) cudaFree(b->slab);
  delete b;
}

static int check_ready(agd_handle* h, agd_batch* b) {
  if (!h || !b) return fail(AGD_ERR_INVALID, "null handle/batch");
  if (!h->dev_weights) return fail(AGD_ERR_INVALID, "weights not loaded (agd_load_weights)");
  return AGD_OK;
}

static int check_overflow(agd_batch* b) {
  int flag[4];
  CUDA_TRY(cudaMemcpy(flag, b->d.counters, sizeof(flag), cudaMemcpyDeviceToHost));
  if (flag[3] != 0) return fail(AGD_ERR_CAPACITY, "a molecule exceeds the adjacency row width chosen for the batch (mol_ptr changed after agd_batch_create?)");
  if ((int64_t)flag[0] > b->d.cap) return fail(AGD_ERR_CAPACITY, "edge count exceeded the declared capacity");
  return AGD_OK;
}

int agd_build_edges(agd_handle* h, agd_batch* b, const float* pos, const agd_forward_out* out, void* stream) {
  int rc = check_ready(h, b);
  if (rc) return rc;
  if (!pos || !out) return fail(AGD_ERR_INVALID, "null argument");
  if ((rc = enter(h, stream))) return rc;
  LaunchCtx c = make_ctx(h);
  launch_build_edges(c, b->d, pos);
  launch_export_edges(c, b->d, *out, false);
  return leave(h, stream);
}

int agd_forward(agd_handle* h, agd_batch* b, const float* pos, const agd_forward_out* out, void* stream) {
  int rc = check_ready(h, b);
  if (rc) return rc;
  if (!pos || !out) return fail(AGD_ERR_INVALID, "null argument");
  if ((rc = enter(h, stream))) return rc;
  LaunchCtx c = make_ctx(h);
  CUDA_TRY(cudaMemsetAsync(b->d.counters + 4, 0, sizeof(int), h->stream));
  run_global_branch(c, b->d, h->w, pos);
  run_local_branch(c, b->d, h->w, pos, nullptr);
  launch_export_edges(c, b->d, *out, true);
  return leave(h, stream);
}

int agd_forward_edges(agd_handle* h, agd_batch* b, const float* pos, const agd_edge_set* es, const agd_forward_out* out, void* stream) {
  int rc = check_ready(h, b);
  if (rc) return rc;
  if (!pos || !es || !out) return fail(AGD_ERR_INVALID, "null argument");
  if (es->n_edges < 0 || es->n_edges > b->d.cap) return fail(AGD_ERR_CAPACITY, "edge set exceeds the batch capacity");
  if ((rc = enter(h, stream))) return rc;
  BatchDev& d = b->d;
  const size_t E = (size_t)es->n_edges, N = (size_t)d.n_atoms;
  auto cp = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream); };
  CUDA_TRY(cp(d.e_src, es->e_src, E * 4)); CUDA_TRY(cp(d.e_dst, es->e_dst, E * 4)); CUDA_TRY(cp(d.e_type, es->e_type, E * 4));
  CUDA_TRY(cp(d.e_canon, es->e_canon, E * 4)); CUDA_TRY(cp(d.e_len, es->e_len, E * 4));
  CUDA_TRY(cp(d.in_ptr, es->in_ptr, (N + 1) * 4)); CUDA_TRY(cp(d.out_ptr, es->out_ptr, (N + 1) * 4));
  CUDA_TRY(cp(d.c_src, es->c_src, E * 4)); CUDA_TRY(cp(d.c_dst, es->c_dst, E * 4)); CUDA_TRY(cp(d.c_type, es->c_type, E * 4));
  CUDA_TRY(cp(d.c_len, es->c_len, E * 4));
  const int n = es->n_edges;
  CUDA_TRY(cudaMemcpyAsync(d.counters, &n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemsetAsync(d.counters + 4, 0, sizeof(int), h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // `n` lives on this stack frame
  LaunchCtx c = make_ctx(h);
  run_global_branch(c, d, h->w, pos, /*build=*/false);
  d.lc_len_in = es->lc_len;
  run_local_branch(c, d, h->w, pos, nullptr);
  d.lc_len_in = nullptr;
  launch_export_edges(c, d, *out, true);
  return leave(h, stream);
}

int agd_sample(agd_handle* h, agd_batch* b, float* pos, const agd_sample_params* p, int32_t* first_nan_step, void* stream) {
  int rc = check_ready(h, b);
  if (rc) return rc;
  if (!pos || !p || p->n_steps < 0) return fail(AGD_ERR_INVALID, "bad sampler arguments");
  if (first_nan_step) *first_nan_step = -1;
  if (p->n_steps == 0) return AGD_OK;
  if ((rc = enter(h, stream))) return rc;
  BatchDev& d = b->d;

  // Everything below runs on h->stream between enter() and leave().  Whatever way the function is left - a CUDA error in the
  // middle of a stream capture included - the guard ends the capture, destroys the instantiated graphs and re-joins the caller's
  // stream, so the handle stays usable.
  struct Guard {
    agd_handle* h;
    void* user_stream;
    cudaGraphExec_t exec_one[2] = {nullptr, nullptr};
    cudaGraphExec_t exec_chunk[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool left = false;
    ~Guard() {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(h->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
        cudaGraph_t g = nullptr;
        cudaStreamEndCapture(h->stream, &g);
        if (g) cudaGraphDestroy(g);
      }
      for (int g = 0; g < 2; ++g) {
        if (exec_one[g]) cudaGraphExecDestroy(exec_one[g]);
        for (int k = 0; k < 2; ++k)
          if (exec_chunk[g][k]) cudaGraphExecDestroy(exec_chunk[g][k]);
      }
      if (!left) {
        cudaEventRecord(h->ev_out, h->stream);
        cudaStreamWaitEvent((cudaStream_t)user_stream, h->ev_out, 0);
        cudaGetLastError();
      }
    }
  } guard{h, stream};

  // per-step schedule -> device
  if (d.sched_cap < p->n_steps) {
    if (d.sched) CUDA_TRY(cudaFree(d.sched));
    d.sched = nullptr;
    d.sched_cap = 0;
    CUDA_TRY(cudaMalloc(&d.sched, sizeof(float) * 4 * (size_t)p->n_steps));
    d.sched_cap = p->n_steps;
  }
  std::vector<float> sched(4 * (size_t)p->n_steps);
  for (int s = 0; s < p->n_steps; ++s) {
    sched[4 * s + 0] = p->sigma[s];
    sched[4 * s + 1] = p->step_size[s];
    sched[4 * s + 2] = p->noise_scale[s];
    sched[4 * s + 3] = p->use_global[s] ? 1.f : 0.f;
  }
  CUDA_TRY(cudaMemcpyAsync(d.sched, sched.data(), sizeof(float) * sched.size(), cudaMemcpyHostToDevice, h->stream));
  const int init[5] = {0, 0, INT_MAX, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(d.counters, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemsetAsync(d.nan_mol, 0x7f, sizeof(int) * (size_t)d.n_mols, h->stream));   // AGD_NAN_NONE = 0x7f7f7f7f
  CUDA_TRY(cudaStreamSynchronize(h->stream));  // `sched`/`init` are stack/heap temporaries

  LaunchCtx c = make_ctx(h);
  StepParams sp;
  sp.w_global = p->w_global; sp.clip = p->clip; sp.clip_local = p->clip_local; sp.clip_pos = p->clip_pos;
  sp.seed = p->seed; sp.noise = p->noise; sp.traj = p->traj; sp.step_offset = p->step_offset;
  auto one_step = [&](bool use_global) {
    if (use_global) run_global_branch(c, d, h->w, pos);
    run_local_branch(c, d, h->w, pos, nullptr);
    sp.use_global = use_global ? 1 : 0;
    Nvtx r("agd.langevin_step");
    launch_step(c, d, pos, sp);
    launch_advance(c, d);
  };
  bool any_global = false, any_local = false;
  for (int s = 0; s < p->n_steps; ++s) (p->use_global[s] ? any_global : any_local) = true;

  // NaN guard (dualenc.py:539-541) without a host sync per step: the launches are issued in windows of POLL; behind every
  // window the first-NaN-step counter is copied to pinned memory and an event is recorded, and before issuing window w + 1 the
  // host waits for the event of window w - 1.  The GPU always has a full window queued, the host is never more than two windows
  // ahead, and a diverged call stops within ~2 windows instead of running all 5000 steps.
  const int POLL = 8;
  h->poll_flag[0] = h->poll_flag[1] = INT_MAX;
  bool nan_seen = false;
  int since_poll = 0, window = 0;
  auto poll = [&]() -> cudaError_t {
    if (++since_poll < POLL) return cudaSuccess;
    since_poll = 0;
    const int w = window & 1;
    cudaError_t e = cudaMemcpyAsync(h->poll_flag + w, d.counters + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventRecord(h->ev_poll[w], h->stream)) != cudaSuccess) return e;
    if (window >= 1) {
      if ((e = cudaEventSynchronize(h->ev_poll[w ^ 1])) != cudaSuccess) return e;
      if (h->poll_flag[w ^ 1] != INT_MAX) nan_seen = true;
    }
    ++window;
    return cudaSuccess;
  };

  if (p->use_cuda_graph) {
    // Per step type (local-only / local+global) two graphs are captured: one step, and a chunk of CHUNK consecutive steps
    // (the device-side step counter makes a multi-step graph valid for any run of equal-type steps).  A graph launch costs
    // ~0.5 ms of front-end time per ~50-node replay that does NOT overlap with the previous replay (measured: replaying a
    // one-step graph is 10 % slower than plain launches), so it is amortised over CHUNK steps; two instantiations of the
    // chunk graph are launched alternately.
    int chunk = 8;
    if (const char* e = std::getenv("AGD_GRAPH_CHUNK")) chunk = std::atoi(e) > 0 ? std::atoi(e) : 1;
    // Which step types replay graphs (A/B switches).  Local-only steps (7 short kernels, 0.7 ms) are launch-bound: graphs win
    // (0.70 vs 0.75 ms per step).  For global steps (45 kernels, 5 ms) it is a wash: 66.6 vs 65.3 conformers/s on the bench batch.
    bool graph_for[2] = {true, true};
    if (const char* e = std::getenv("AGD_GRAPH_GLOBAL")) graph_for[1] = (e[0] != '0');
    if (const char* e = std::getenv("AGD_GRAPH_LOCAL")) graph_for[0] = (e[0] != '0');
    int64_t per_step_launches[2] = {0, 0};
    // longest run of equal-type steps decides whether a chunk graph is worth building
    int longest[2] = {0, 0};
    for (int s = 0; s < p->n_steps;) {
      int e = s;
      while (e < p->n_steps && (p->use_global[e] != 0) == (p->use_global[s] != 0)) ++e;
      const int g = p->use_global[s] ? 1 : 0;
      if (e - s > longest[g]) longest[g] = e - s;
      s = e;
    }
    for (int g = 0; g < 2; ++g) {
      if (!(g ? any_global : any_local) || !graph_for[g]) continue;
      for (int kind = 0; kind < 2; ++kind) {   // 0: one step, 1: chunk
        if (kind == 1 && (chunk < 2 || longest[g] < chunk)) continue;
        cudaGraph_t graph = nullptr;
        const int64_t before = h->launches;
        CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        for (int r = 0; r < (kind ? chunk : 1); ++r) one_step(g == 1);
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (kind == 0) per_step_launches[g] = h->launches - before;
        h->launches = before;
        if (e != cudaSuccess) return fail(AGD_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        if (kind == 0) e = cudaGraphInstantiate(&guard.exec_one[g], graph, 0);
        else
          for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaGraphInstantiate(&guard.exec_chunk[g][k], graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return fail(AGD_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
      }
    }
    cudaError_t e = cudaSuccess;
    int flip = 0;
    for (int s = 0; s < p->n_steps && e == cudaSuccess && !nan_seen;) {
      const int g = p->use_global[s] ? 1 : 0;
      if (!graph_for[g]) {
        one_step(g == 1);
        s += 1;
        e = poll();
        continue;
      }
      int run = 1;
      while (run < chunk && s + run < p->n_steps && (p->use_global[s + run] != 0) == (g == 1)) ++run;
      if (run == chunk && guard.exec_chunk[g][0]) {
        e = cudaGraphLaunch(guard.exec_chunk[g][flip], h->stream);
        flip ^= 1;
        h->launches += per_step_launches[g] * chunk;
        s += chunk;
      } else {
        e = cudaGraphLaunch(guard.exec_one[g], h->stream);
        h->launches += per_step_launches[g];
        s += 1;
      }
      if (e == cudaSuccess) e = poll();
    }
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(AGD_ERR_CUDA, std::string("graph launch: ") + cudaGetErrorString(e));
    if (e2 != cudaSuccess) return fail(AGD_ERR_CUDA, std::string("sampling loop: ") + cudaGetErrorString(e2));
  } else {
    for (int s = 0; s < p->n_steps && !nan_seen; ++s) {
      one_step(p->use_global[s] != 0);
      CUDA_TRY(poll());
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  CUDA_TRY(cudaGetLastError());
  if ((rc = check_overflow(b))) return rc;
  int range_flag = 0;
  CUDA_TRY(cudaMemcpy(&range_flag, d.counters + 4, sizeof(int), cudaMemcpyDeviceToHost));
  int nan_step = INT_MAX;
  CUDA_TRY(cudaMemcpy(&nan_step, d.counters + 2, sizeof(int), cudaMemcpyDeviceToHost));
  guard.left = true;
  if ((rc = leave(h, stream))) return rc;
  if (range_flag) return fail(AGD_ERR_RANGE, "an activation left the fp16-split range; re-run in AGD_MODE_TF32");
  if (nan_step != INT_MAX) {
    if (first_nan_step) *first_nan_step = nan_step;
    return fail(AGD_ERR_NAN, "NaN positions at sampling step " + std::to_string(nan_step));
  }
  return AGD_OK;
}

int agd_nan_steps(agd_batch* b, int32_t* host_out, int32_t n_mols) {
  if (!b || !host_out || n_mols != b->d.n_mols) return fail(AGD_ERR_INVALID, "agd_nan_steps: bad arguments");
  CUDA_TRY(cudaSetDevice(b->h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->h->stream));
  CUDA_TRY(cudaMemcpy(host_out, b->d.nan_mol, sizeof(int) * (size_t)n_mols, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n_mols; ++i)
    if (host_out[i] == 0x7f7f7f7f) host_out[i] = -1;
  return AGD_OK;
}

int agd_extend_bond_order(const int32_t* mol_ptr, int32_t n_mols, int32_t n_atoms, const int32_t* bond_ptr,
                          const int32_t* bond_dst, const int32_t* bond_type, int32_t order, int32_t num_bond_types,
                          int32_t* out_count, const int32_t* out_ptr, int32_t* out_dst, int32_t* out_type, void* stream) {
  if (!mol_ptr || !bond_ptr || n_mols <= 0) return fail(AGD_ERR_INVALID, "bad arguments");
  if (!out_ptr && !out_count) return fail(AGD_ERR_INVALID, "count pass needs out_count");
  const int mw = adjacency_words(mol_ptr, n_mols);
  if (mw == -2) return fail(AGD_ERR_CUDA, "cannot read mol_ptr (device pointer expected)");
  if (mw < 0) return fail(AGD_ERR_CAPACITY, "a molecule exceeds AGD_MAX_MOL_ATOMS");
  int rc = launch_extend_bond_order((cudaStream_t)stream, mol_ptr, n_mols, n_atoms, bond_ptr, bond_dst, bond_type, order,
                                    num_bond_types, out_count, out_ptr, out_dst, out_type, mw);
  if (rc) return fail(rc, "edge order must be 1..3");
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

int agd_op_cfconv_aggregate(const float* x, const float* W, const int32_t* src, const int32_t* in_ptr, int32_t n_nodes,
                            int32_t F, float* out, void* stream) {
  if (F != 64 && F != 128 && F != 192) return fail(AGD_ERR_INVALID, "F must be 64, 128 or 192");
  int64_t dummy = 0;
  LaunchCtx c{};
  c.stream = (cudaStream_t)stream;
  c.launch_counter = &dummy;
  c.prof = nullptr;
  c.use_tc = 0;
  c.f16_fuse = 0;
  c.f16_mlp = 0;
  c.f16_pair = 0;
  c.f16_node = 0;
  c.local_pairs = 0;
  c.f16_debug_filt = 0;
  c.f16_timing = nullptr;
  c.mlp_act = 0;
  launch_aggregate(c, x, W, src, in_ptr, n_nodes, F, out);
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

int agd_op_eq_transform(const float* score, const float* pos, const int32_t* src, const int32_t* dst, const float* length,
                        int64_t n_edges, int32_t n_nodes, float* out, void* stream) {
  launch_eq_transform((cudaStream_t)stream, score, pos, src, dst, length, n_edges, n_nodes, out);
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

int agd_op_gin_message(const float* x, const float* ea, const int32_t* src, const int32_t* in_ptr, int32_t n_nodes, float eps,
                       float* out, void* stream) {
  if (!x || !ea || !src || !in_ptr || !out || n_nodes < 0) return fail(AGD_ERR_INVALID, "bad arguments");
  launch_gin_message((cudaStream_t)stream, x, ea, nullptr, src, in_ptr, n_nodes, eps, nullptr, out);
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

int agd_op_eq_transform_segments(const float* pos, const float* score_out, const int32_t* col_of_out, const int32_t* out_ptr,
                                 const float* score_in, const int32_t* row_of_in, const int32_t* in_ptr, int32_t n_nodes,
                                 float* out, void* stream) {
  if (!pos || !score_out || !col_of_out || !out_ptr || !score_in || !row_of_in || !in_ptr || !out || n_nodes < 0)
    return fail(AGD_ERR_INVALID, "bad arguments");
  launch_eq_transform_segments((cudaStream_t)stream, pos, score_out, col_of_out, out_ptr, score_in, row_of_in, in_ptr, n_nodes, out);
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

int agd_op_kabsch_rmsd(const float* ref, const float* gen, const int32_t* sel, int32_t n_sel, int32_t n_atoms, int32_t n_ref,
                       int32_t n_gen, float* out, void* stream) {
  if (!ref || !gen || !out || n_atoms <= 0 || n_ref < 0 || n_gen < 0) return fail(AGD_ERR_INVALID, "bad arguments");
  if (!sel) n_sel = n_atoms;
  if (n_sel <= 0 || n_sel > n_atoms) return fail(AGD_ERR_INVALID, "atom selection must hold 1..n_atoms indices");
  launch_kabsch_rmsd((cudaStream_t)stream, ref, gen, sel, n_sel, n_atoms, n_ref, n_gen, out);
  CUDA_TRY(cudaGetLastError());
  return AGD_OK;
}

int64_t agd_debug_fetch(agd_batch* b, const char* name, float* dst, int64_t capacity) {
  if (!b || !name || !dst) return fail(AGD_ERR_INVALID, "null argument");
  const BatchDev& d = b->d;
  cudaSetDevice(b->h->cfg.device);
  cudaStreamSynchronize(b->h->stream);
  int n_edges = 0;
  cudaMemcpy(&n_edges, d.counters, sizeof(int), cudaMemcpyDeviceToHost);
  const float* src = nullptr;
  int64_t n = 0;
  const std::string s(name);
  if (s == "g2") { src = d.g2; n = (int64_t)n_edges * HID; }
  else if (s == "filt") { src = d.filt; n = (int64_t)n_edges * 192; }
  else if (s == "h_global") { src = d.h; n = (int64_t)d.n_atoms * HID; }
  else if (s == "xcat") { src = d.xcat; n = (int64_t)d.n_atoms * 192; }
  else if (s == "agg") { src = d.agg; n = (int64_t)d.n_atoms * 192; }
  else if (s == "ea_local") {
    n = (int64_t)d.n_local * HID;
    if (d.pairs_used) {   // ea_loc holds one row per pair: hand out the rows in local-edge order
      if (n > capacity) return fail(AGD_ERR_CAPACITY, "destination too small");
      launch_gather_rows128(b->h->stream, d.ea_loc, d.lp_of, d.n_local, dst);
      if (cudaStreamSynchronize(b->h->stream) != cudaSuccess) return fail(AGD_ERR_CUDA, "gather failed");
      return n;
    }
    src = d.ea_loc;
  }
  else if (s == "h_local") { src = (b->h->cfg.num_convs_local % 2) ? d.gx1 : d.gx0; n = (int64_t)d.n_atoms * HID; }
  else if (s == "e_len") { src = d.e_len; n = n_edges; }
  else if (s == "s_csc") { src = d.s_csc; n = n_edges; }
  else return fail(AGD_ERR_INVALID, "unknown tensor name");
  if (n > capacity) return fail(AGD_ERR_CAPACITY, "destination too small");
  cudaError_t e = cudaMemcpy(dst, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) return fail(AGD_ERR_CUDA, cudaGetErrorString(e));
  return n;
}

int64_t agd_launch_count(const agd_handle* h) { return h ? h->launches : 0; }

int agd_set_mode(agd_handle* h, int mode) {
  if (!h || mode < AGD_MODE_FFMA || mode > AGD_MODE_F16) return fail(AGD_ERR_INVALID, "mode must be 0, 1 or 2");
  h->use_tc = mode;
  return AGD_OK;
}
int agd_get_mode(const agd_handle* h) { return h ? h->use_tc : AGD_ERR_INVALID; }
int agd_set_option(agd_handle* h, const char* name, int value) {
  if (!h || !name) return fail(AGD_ERR_INVALID, "null argument");
  if (std::strcmp(name, "f16_fuse") == 0) h->f16_fuse = value ? 1 : 0;
  else if (std::strcmp(name, "f16_mlp") == 0) h->f16_mlp = value ? 1 : 0;
  else if (std::strcmp(name, "f16_pair") == 0) h->f16_pair = value ? 1 : 0;
  else if (std::strcmp(name, "f16_node") == 0) h->f16_node = value ? 1 : 0;
  else if (std::strcmp(name, "local_pairs") == 0) h->local_pairs = value ? 1 : 0;
  else if (std::strcmp(name, "mlp_act") == 0) {
    if (value < 0 || value >= AGD_ACT_COUNT) return fail(AGD_ERR_INVALID, "unknown mlp_act id");
    h->mlp_act = value;
  }
  else if (std::strcmp(name, "f16_debug_filt") == 0) h->f16_debug_filt = value;   // bit 0: write filt; bits 1.. : timing experiments (tc_cfconv.cu)
  else if (std::strcmp(name, "f16_timing") == 0) {   // diagnostics: 1 = allocate + zero the phase counters, 0 = off
    if (value && !h->f16_timing) {
      if (cudaMalloc(&h->f16_timing, 64 * sizeof(unsigned long long)) != cudaSuccess) return fail(AGD_ERR_CUDA, "cudaMalloc");
    }
    if (value) cudaMemset(h->f16_timing, 0, 64 * sizeof(unsigned long long));
    else if (h->f16_timing) { cudaFree(h->f16_timing); h->f16_timing = nullptr; }
  }
  else return fail(AGD_ERR_INVALID, std::string("unknown option ") + name);
  return AGD_OK;
}
int agd_f16_lo_shift(void) { return f16_lo_shift(); }
int agd_debug_timing(agd_handle* h, uint64_t* out64) {
  if (!h || !out64 || !h->f16_timing) return fail(AGD_ERR_INVALID, "timing is off (agd_set_option f16_timing 1)");
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemcpy(out64, h->f16_timing, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return AGD_OK;
}
int agd_range_flag(agd_batch* b, int32_t* flag_out) {
  if (!b || !flag_out) return fail(AGD_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(b->h->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(b->h->stream));
  CUDA_TRY(cudaMemcpy(flag_out, b->d.counters + 4, sizeof(int), cudaMemcpyDeviceToHost));
  return AGD_OK;
}

int agd_profile_forward(agd_handle* h, agd_batch* b, const float* pos, int32_t with_global, char* labels, int64_t labels_cap,
                        float* ms, int32_t cap, int32_t* n_out, int32_t* n_edges_out, void* stream) {
  int rc = check_ready(h, b);
  if (rc) return rc;
  if (!pos || !labels || !ms || !n_out) return fail(AGD_ERR_INVALID, "null argument");
  if ((rc = enter(h, stream))) return rc;
  LaunchCtx c = make_ctx(h);
  Prof prof;
  cudaEvent_t start;
  CUDA_TRY(cudaEventCreate(&start));
  CUDA_TRY(cudaEventRecord(start, h->stream));
  c.prof = &prof;
  if (with_global) run_global_branch(c, b->d, h->w, pos);
  run_local_branch(c, b->d, h->w, pos, nullptr);
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  std::string names;
  int n = 0;
  cudaEvent_t prev = start;
  for (size_t i = 0; i < prof.ev.size(); ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, prev, prof.ev[i]);
    if (n < cap) {
      ms[n++] = t;
      names += prof.label[i];
      names += '\n';
    }
    prev = prof.ev[i];
  }
  cudaEventDestroy(start);
  for (auto e : prof.ev) cudaEventDestroy(e);
  if ((int64_t)names.size() + 1 > labels_cap) return fail(AGD_ERR_CAPACITY, "label buffer too small");
  std::memcpy(labels, names.c_str(), names.size() + 1);
  *n_out = n;
  if (n_edges_out) CUDA_TRY(cudaMemcpy(n_edges_out, b->d.counters, sizeof(int), cudaMemcpyDeviceToHost));
  return leave(h, stream);
}

}  // extern "C"
