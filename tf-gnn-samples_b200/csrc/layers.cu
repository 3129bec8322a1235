// layers.cu -- C-ABI entry points: one forward per reference layer function (include/rgnn.h).
//
// Every layer is re-associated "transform first" (SURVEY.md 7): the per-type Dense is applied to the
// V node rows (one tensor-core GEMM over all types, T = H . [W_0|...|W_{L-1}]) instead of to the M
// gathered edge rows (gnns/rgcn.py:88,98 does M*D*D*2 FLOP; this does V*L*D*D*2), and the edge stage
// is one fused sorted-segment kernel (seg_kernels.cu).  RGAT already works this way in the reference
// (rgat.py:95-96).  Only an MLP's layers after a per-edge nonlinearity stay per-edge (edge-MLP with
// >= 1 hidden layer and target input): those run as per-type row-range GEMMs over materialised rows.
#include <mutex>
#include <atomic>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "plan.cuh"
#include "seg.cuh"

namespace rgnn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  size_t cap, used = 0;
  bool overflow = false;
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c) {}
  float* floats(size_t n) {
    const size_t bytes = align_up(n * sizeof(float), 256);
    if (base == nullptr || used + bytes > cap) { overflow = true; used += bytes; return nullptr; }
    float* p = reinterpret_cast<float*>(base + used);
    used += bytes;
    return p;
  }
};

int check_common(const rgnn_plan_t* plan, const float* h, int d_in, int d_out, const float* out, int num_timesteps,
                 const char* who) {
  RGNN_REQUIRE(plan != nullptr, "%s: plan is NULL", who);
  RGNN_REQUIRE(h != nullptr && out != nullptr, "%s: node_embeddings / out is NULL", who);
  RGNN_REQUIRE(h != out, "%s: out must not alias node_embeddings", who);
  RGNN_REQUIRE(aligned16(h) && aligned16(out), "%s: node_embeddings / out must be 16-byte aligned", who);
  RGNN_REQUIRE(d_in > 0 && d_out > 0 && (d_in % 4) == 0 && (d_out % 4) == 0,
               "%s: state dims must be positive multiples of 4 (d_in=%d, d_out=%d)", who, d_in, d_out);
  RGNN_REQUIRE(num_timesteps >= 1, "%s: num_timesteps %d < 1", who, num_timesteps);
  RGNN_REQUIRE(num_timesteps == 1 || plan->Vt == plan->V,
               "%s: a plan restricted to %d of %d target rows supports num_timesteps == 1 only (halo rows are not updated)", who, plan->Vt, plan->V);
  RGNN_REQUIRE(num_timesteps == 1 || d_in == d_out,
               "%s: num_timesteps > 1 needs state_dim == input dim (d_in=%d, d_out=%d)", who, d_in, d_out);
  return RGNN_OK;
}
int check_act(int act, const char* who) {
  RGNN_REQUIRE(act >= RGNN_ACT_LINEAR && act <= RGNN_ACT_GELU, "%s: Unknown activation function code %d", who, act);
  return RGNN_OK;
}
int check_agg(int agg, const char* who) {
  RGNN_REQUIRE(agg >= RGNN_AGG_SUM && agg <= RGNN_AGG_SQRT_N, "%s: Unknown aggregation function code %d", who, agg);
  return RGNN_OK;
}
int check_ws(const Arena& a, const char* who) {
  if (a.overflow) {
    set_error("%s: workspace too small (%zu bytes given, %zu needed)", who, a.cap, a.used);
    return RGNN_E_WORKSPACE;
  }
  return RGNN_OK;
}

// Dispatch one dense contraction on the tcgen05 kernel (weight images packed into arena scratch that is released
// right after the enqueue -- later users are stream-ordered).
int run_gemm(const GemmParams& g, Arena& ar, cudaStream_t stream) {
  const size_t need = gemm_tc_pack_bytes(g);
  const size_t mark = ar.used;
  void* ws = ar.floats(need / sizeof(float));
  if (ar.overflow) {
    set_error("workspace too small for the weight images of a dense contraction (%zu bytes given, %zu needed)", ar.cap, ar.used);
    return RGNN_E_WORKSPACE;
  }
  const int rc = launch_gemm_tcgen05(g, ws, need, stream);
  ar.used = mark;
  return rc;
}

// T[V, batch*N] = A[V, K] . B_z  for z < batch  (shared A)
int gemm_shared_a(Arena& ar, const float* A, int V, int K, const float* const* B, int batch, int ldb, int N, float* C, int act,
                  cudaStream_t stream) {
  GemmParams g;
  g.A1 = A; g.lda1 = K; g.K1 = K;
  g.M = V; g.N = N;
  g.C = C; g.ldc = batch * N;
  g.ldb1 = ldb;
  g.act = act;
  g.batch_mode = BATCH_SHARED_A; g.batch = batch;
  for (int z = 0; z < batch; ++z) { g.bptr[z] = B[z]; g.bptr2[z] = nullptr; }
  return run_gemm(g, ar, stream);
}

// The per-type Dense on the SOURCE side of the edge stage (rgcn.py:98, ggnn.py:80-82, gnn_film.py:94 applied to nodes).
// Dense form: T[V, L, D] = cur . [W_0|..|W_{L-1}].  Sparsely typed graphs (the plan holds a compact pair table, plan.cuh):
// only the (source, type) rows that some edge gathers are transformed -- a row-range GEMM per type whose A rows follow the
// pair list -- and the edge stage addresses the compact table.  Call after seg_from_plan(s, plan).
int transform_sources(const rgnn_plan_t* plan, Arena& ar, const float* cur, int d_in, int D, const float* const* W, float* T,
                      cudaStream_t stream, SegParams& s) {
  const int V = plan->V, L = plan->L;
  RGNN_PROPAGATE(plan_wait_sources(plan, stream));   // an overlapped halo exchange must have landed before source rows are read
  if (plan->n_pairs >= 0 && plan->pair_src != nullptr) {
    GemmParams g;
    g.A1 = cur; g.lda1 = d_in; g.K1 = d_in; g.a_rows = plan->pair_src;
    g.M = plan->n_pairs; g.N = D; g.C = T; g.ldc = D; g.ldb1 = D;
    g.batch_mode = BATCH_ROW_RANGES; g.batch = L; g.max_rows = plan->max_type_pairs;
    for (int l = 0; l < L; ++l) { g.bptr[l] = W[l]; g.bptr2[l] = nullptr; g.row_off[l] = plan->pair_type_off[l]; }
    g.row_off[L] = plan->pair_type_off[L];
    RGNN_PROPAGATE(run_gemm(g, ar, stream));
    s.table = T; s.e_idx = plan->e_pair; s.stride_idx = D; s.stride_type = 0;
    return RGNN_OK;
  }
  RGNN_PROPAGATE(gemm_shared_a(ar, cur, V, d_in, W, L, D, D, T, RGNN_ACT_LINEAR, stream));
  s.table = T; s.stride_idx = (long)L * D; s.stride_type = D;
  return RGNN_OK;
}

// scratch for the multi-CTA split of heavy targets (seg_kernels.cu); nothing when the plan is known to have none
size_t heavy_scratch_floats(const rgnn_plan_t* plan, size_t d) {
  return plan->num_heavy_host == 0 ? 0 : (size_t)plan->heavy_items_cap * d;
}
void seg_heavy_scratch(SegParams& s, const rgnn_plan_t* plan, Arena& ar, int d) {
  if (plan->num_heavy_host == 0 || plan->heavy_items == nullptr) return;
  float* scratch = ar.floats((size_t)plan->heavy_items_cap * d);
  if (ar.overflow) return;   // the caller's check_ws reports it
  s.heavy_scratch = scratch;
}

void seg_from_plan(SegParams& s, const rgnn_plan_t* plan) {
  s.V = plan->Vt; s.L = plan->L; s.scale_ld = plan->V;   // only the wanted target rows are reduced (rgnn_plan_set_num_targets)
  s.heavy_list = plan->heavy_list; s.heavy_count = plan->err_flag + 1;
  s.heavy_base = plan->heavy_base; s.heavy_items = plan->heavy_items; s.heavy_item_count = plan->err_flag + 3;
  s.heavy_items_known = plan->num_heavy_items_host; s.heavy_items_cap = plan->heavy_items_cap; s.heavy_chunk = RGNN_HEAVY_CHUNK;
  s.heavy_threshold = RGNN_HEAVY_SEGMENT; s.heavy_known = plan->num_heavy_host;
  s.seg_off = plan->seg_off; s.e_type = plan->e_type; s.e_idx = plan->e_src;
}

// Where the per-message rows of an MLP-style layer live after the dense stages.
struct MsgSource {
  const float* table = nullptr;
  const int32_t* idx = nullptr;
  long stride_idx = 0, stride_type = 0;
  int width = 0;
  int msg_mode = MSG_LINEAR;
  const float* mod_table = nullptr;
  long mod_sn = 0, mod_st = 0;
};

// Evaluate MLP_l([h_u | h_v]) / MLP_l(h_u) for every message (gnn_edge_mlp.py:87-102, rgin.py:106-124).
// kernels: type-major [L][nl]; dims [nl+1]; nl = number of Dense kernels (hidden layers + 1), 0 = no MLP.
int build_mlp_messages(const rgnn_plan_t* plan, const float* cur, int d_in, const float* const* kernels,
                       const int32_t* dims, int nl, int use_target, int hidden_act, Arena& ar, cudaStream_t stream,
                       MsgSource* out) {
  const int V = plan->V, L = plan->L;
  const int M = (int)plan->M;
  MsgSource ms;
  if (nl == 0) {
    if (!use_target) {
      ms.table = cur; ms.idx = plan->e_src; ms.stride_idx = d_in; ms.stride_type = 0; ms.width = d_in;
    } else {   // messages are the raw [h_u | h_v] pairs (rgin.py:114-124 with edge MLP None)
      float* X = ar.floats((size_t)M * 2 * d_in);
      if (ar.overflow) { *out = ms; return RGNN_OK; }
      EdgeBuildParams b;
      b.L = L; b.D = d_in; b.o_src = plan->o_src; b.o_tgt = plan->o_tgt;
      memcpy(b.type_off, plan->type_off, sizeof(b.type_off)); b.max_type_edges = plan->max_type_edges;
      b.p = cur; b.p_stride_node = d_in; b.p_stride_type = 0; b.concat = 1;
      b.x = X; b.ldx = 2 * d_in;
      RGNN_PROPAGATE(launch_edge_build(b, stream));
      ms.table = X; ms.idx = plan->e_orig; ms.stride_idx = 2 * d_in; ms.stride_type = 0; ms.width = 2 * d_in;
    }
    *out = ms;
    return RGNN_OK;
  }
  RGNN_REQUIRE(nl <= RGNN_MAX_MLP_LAYERS, "edge MLP with %d layers exceeds the supported %d", nl, RGNN_MAX_MLP_LAYERS);
  RGNN_REQUIRE(dims[0] == d_in * (use_target ? 2 : 1), "edge MLP input dim %d does not match %d", dims[0],
               d_in * (use_target ? 2 : 1));
  for (int j = 1; j <= nl; ++j) RGNN_REQUIRE(dims[j] > 0 && (dims[j] % 4) == 0, "edge MLP dim %d must be a positive multiple of 4", dims[j]);
  for (int i = 0; i < L * nl; ++i) RGNN_REQUIRE(kernels[i] != nullptr, "edge MLP kernel %d is NULL", i);

  const float* bp[RGNN_MAX_EDGE_TYPES];
  if (!use_target) {
    // whole MLP is per (node, type): chain of node-level GEMMs
    float* prev = ar.floats((size_t)V * L * dims[1]);
    if (ar.overflow) { *out = ms; return RGNN_OK; }
    for (int l = 0; l < L; ++l) bp[l] = kernels[l * nl + 0];
    RGNN_PROPAGATE(gemm_shared_a(ar, cur, V, d_in, bp, L, dims[1], dims[1], prev, nl > 1 ? hidden_act : RGNN_ACT_LINEAR, stream));
    for (int j = 1; j < nl; ++j) {
      float* next = ar.floats((size_t)V * L * dims[j + 1]);
      if (ar.overflow) { *out = ms; return RGNN_OK; }
      GemmParams g;
      g.A1 = prev; g.lda1 = L * dims[j]; g.K1 = dims[j];
      g.M = V; g.N = dims[j + 1]; g.C = next; g.ldc = L * dims[j + 1]; g.ldb1 = dims[j + 1];
      g.act = (j < nl - 1) ? hidden_act : RGNN_ACT_LINEAR;
      g.batch_mode = BATCH_COL_BLOCKS; g.batch = L;
      for (int l = 0; l < L; ++l) { g.bptr[l] = kernels[l * nl + j]; g.bptr2[l] = nullptr; }
      RGNN_PROPAGATE(run_gemm(g, ar, stream));
      prev = next;
    }
    ms.table = prev; ms.idx = plan->e_src; ms.stride_idx = (long)L * dims[nl]; ms.stride_type = dims[nl]; ms.width = dims[nl];
    *out = ms;
    return RGNN_OK;
  }
  // use_target: first Dense splits into a source half P and a target half Q of the kernel rows
  RGNN_REQUIRE(2 * L <= RGNN_MAX_EDGE_TYPES, "edge MLP with target input supports at most %d edge types", RGNN_MAX_EDGE_TYPES / 2);
  const int d1 = dims[1];
  float* PQ = ar.floats((size_t)V * 2 * L * d1);
  if (ar.overflow) { *out = ms; return RGNN_OK; }
  for (int l = 0; l < L; ++l) {
    bp[l] = kernels[l * nl + 0];                              // rows [0, d_in)      multiply h_u
    bp[L + l] = kernels[l * nl + 0] + (size_t)d_in * d1;      // rows [d_in, 2 d_in) multiply h_v
  }
  RGNN_PROPAGATE(gemm_shared_a(ar, cur, V, d_in, bp, 2 * L, d1, d1, PQ, RGNN_ACT_LINEAR, stream));
  if (nl == 1) {
    ms.table = PQ; ms.idx = plan->e_src; ms.stride_idx = 2L * L * d1; ms.stride_type = d1; ms.width = d1;
    ms.msg_mode = MSG_ADDTGT; ms.mod_table = PQ + (size_t)L * d1; ms.mod_sn = 2L * L * d1; ms.mod_st = d1;
    *out = ms;
    return RGNN_OK;
  }
  float* X = ar.floats((size_t)M * d1);
  if (ar.overflow) { *out = ms; return RGNN_OK; }
  {
    EdgeBuildParams b;
    b.L = L; b.D = d1; b.o_src = plan->o_src; b.o_tgt = plan->o_tgt;
    memcpy(b.type_off, plan->type_off, sizeof(b.type_off)); b.max_type_edges = plan->max_type_edges;
    b.p = PQ; b.p_stride_node = 2L * L * d1; b.p_stride_type = d1;
    b.q = PQ + (size_t)L * d1; b.q_stride_node = 2L * L * d1; b.q_stride_type = d1;
    b.act = hidden_act; b.x = X; b.ldx = d1;
    RGNN_PROPAGATE(launch_edge_build(b, stream));
  }
  float* prev = X;
  for (int j = 1; j < nl; ++j) {
    float* next = ar.floats((size_t)M * dims[j + 1]);
    if (ar.overflow) { *out = ms; return RGNN_OK; }
    GemmParams g;
    g.A1 = prev; g.lda1 = dims[j]; g.K1 = dims[j];
    g.M = M; g.N = dims[j + 1]; g.C = next; g.ldc = dims[j + 1]; g.ldb1 = dims[j + 1];
    g.act = (j < nl - 1) ? hidden_act : RGNN_ACT_LINEAR;
    g.batch_mode = BATCH_ROW_RANGES; g.batch = L; g.max_rows = plan->max_type_edges;
    for (int l = 0; l < L; ++l) { g.bptr[l] = kernels[l * nl + j]; g.bptr2[l] = nullptr; g.row_off[l] = plan->type_off[l]; }
    g.row_off[L] = plan->type_off[L];
    RGNN_PROPAGATE(run_gemm(g, ar, stream));
    prev = next;
  }
  ms.table = prev; ms.idx = plan->e_orig; ms.stride_idx = dims[nl]; ms.stride_type = 0; ms.width = dims[nl];
  *out = ms;
  return RGNN_OK;
}

void seg_from_source(SegParams& s, const MsgSource& ms) {
  s.table = ms.table; s.e_idx = ms.idx; s.stride_idx = ms.stride_idx; s.stride_type = ms.stride_type;
  s.D = ms.width; s.msg_mode = ms.msg_mode; s.mod_table = ms.mod_table;
  s.mod_stride_node = ms.mod_sn; s.mod_stride_type = ms.mod_st;
}

}  // namespace
}  // namespace rgnn

using namespace rgnn;

extern "C" int rgnn_version(void) { return RGNN_VERSION; }
extern "C" const char* rgnn_last_error(void) { return g_err; }
extern "C" int64_t rgnn_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int rgnn_set_weight_cache(int enable) {
  gemm_weight_cache_enable(enable != 0);
  if (!enable) gemm_weight_cache_clear();
  return RGNN_OK;
}
extern "C" int rgnn_weight_cache_clear(void) {
  gemm_weight_cache_clear();
  return RGNN_OK;
}

extern "C" size_t rgnn_workspace_bytes(const rgnn_plan_t* plan, int layer_kind, int32_t d_in, int32_t d_out,
                                       int32_t mlp_layers) {
  if (plan == nullptr || d_in <= 0 || d_out <= 0) return 0;
  const size_t V = (size_t)plan->V, L = (size_t)plan->L, M = (size_t)plan->M;
  const size_t dm = (size_t)((2 * d_in > d_out) ? 2 * d_in : d_out);
  const size_t pad = 256 * 32;
  size_t floats = 0;
  switch (layer_kind) {
    case RGNN_LAYER_RGCN: floats = V * 2 * L * dm + 2 * V * dm; break;
    case RGNN_LAYER_GGNN: floats = V * L * dm + 6 * V * dm; break;
    case RGNN_LAYER_RGAT: floats = V * L * dm + 2 * V * L * dm / 4 + 2 * V * dm; break;
    case RGNN_LAYER_FILM: floats = 3 * V * L * dm + 2 * V * dm; break;
    case RGNN_LAYER_RGCN_BACKWARD: floats = 2 * V * L * dm + 2 * V * dm + (148 * 16384 + L * dm * dm) + 64 * 1024; break;   // + split-K partial tiles of dW
    case RGNN_LAYER_EDGE_MLP:
    case RGNN_LAYER_RGIN: {
      const size_t nl = (size_t)(mlp_layers > 0 ? mlp_layers : 1);
      floats = V * 2 * L * dm * (nl + 1) + M * dm * (nl + 1) + (4 + nl) * V * dm;
      break;
    }
    case RGNN_LAYER_RGDCN: floats = V * L * dm * (size_t)(mlp_layers > 0 ? mlp_layers : 16) + 2 * V * dm; break;   // mlp_layers carries channel_dim
    default: return 0;
  }
  // scratch for the pre-swizzled hi/lo weight images of the largest dense contraction of the layer
  const size_t pack = 2 * (2 * dm + 64) * (2 * L * dm + 2 * dm + 512);
  floats += heavy_scratch_floats(plan, dm) + 64;   // partial rows of split heavy targets
  return (floats + pack) * sizeof(float) + pad;
}

// ---------------------------------------------------------------------------------------------
// gnns/rgcn.py:8-117
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_rgcn_forward(const rgnn_plan_t* plan, const float* h, int32_t d_in, int32_t d_out,
                                 const float* const* edge_weights, const float* num_incoming, int activation,
                                 int aggregation, int normalize, int both, int num_timesteps, float* out,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d_in, d_out, out, num_timesteps, "rgcn"));
  RGNN_PROPAGATE(plan_wait_sources(plan, stream));
  RGNN_PROPAGATE(check_act(activation, "rgcn"));
  RGNN_PROPAGATE(check_agg(aggregation, "rgcn"));
  RGNN_REQUIRE(edge_weights != nullptr, "rgcn: edge_weights is NULL");
  RGNN_REQUIRE(!normalize || num_incoming != nullptr, "rgcn: normalize_by_num_incoming needs type_to_num_incoming_edges");
  const int V = plan->V, L = plan->L;
  RGNN_REQUIRE(!both || 2 * L <= RGNN_MAX_EDGE_TYPES, "rgcn: use_both_source_and_target supports at most %d edge types", RGNN_MAX_EDGE_TYPES / 2);
  for (int l = 0; l < L; ++l) RGNN_REQUIRE(edge_weights[l] != nullptr, "rgcn: edge weight %d is NULL", l);
  Arena ar(workspace, workspace_bytes);
  const int nb = both ? 2 * L : L;
  float* T = ar.floats((size_t)V * nb * d_out);
  float* buf[2] = {nullptr, nullptr};
  if (num_timesteps > 1) { buf[0] = ar.floats((size_t)V * d_out); buf[1] = ar.floats((size_t)V * d_out); }
  SegParams heavy;
  seg_heavy_scratch(heavy, plan, ar, d_out);
  RGNN_PROPAGATE(check_ws(ar, "rgcn"));

  const float* bp[RGNN_MAX_EDGE_TYPES];
  const float* cur = h;
  int din = d_in;
  for (int t = 0; t < num_timesteps; ++t) {                                   // rgcn.py:81
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    for (int l = 0; l < L; ++l) {
      bp[l] = edge_weights[l];                                                // kernel rows [0, d_in): source half
      if (both) bp[L + l] = edge_weights[l] + (size_t)din * d_out;            // rows [d_in, 2 d_in): target half (rgcn.py:95)
    }
    SegParams s;
    seg_from_plan(s, plan);
    s.D = d_out;
    if (both) {
      RGNN_PROPAGATE(gemm_shared_a(ar, cur, V, din, bp, nb, d_out, d_out, T, RGNN_ACT_LINEAR, stream));   // rgcn.py:98 on nodes
      s.table = T; s.stride_idx = (long)nb * d_out; s.stride_type = d_out;
    } else {
      RGNN_PROPAGATE(transform_sources(plan, ar, cur, din, d_out, bp, T, stream, s));                     // rgcn.py:98 on the used (source, type) rows
    }
    s.num_incoming = normalize ? num_incoming : nullptr;                      // rgcn.py:100-104
    if (both) { s.msg_mode = MSG_ADDTGT; s.mod_table = T + (size_t)L * d_out; s.mod_stride_node = (long)nb * d_out; s.mod_stride_type = d_out; }
    s.agg = aggregation; s.act_out = activation;                              // rgcn.py:110,114
    s.out = dst; s.ld_out = d_out; s.heavy_scratch = heavy.heavy_scratch;
    RGNN_PROPAGATE(launch_seg_reduce(s, stream));
    cur = dst; din = d_out;
  }
  return RGNN_OK;
}

// Backward of ONE timestep of sparse_rgcn_layer (source-only messages): what tf.gradients produces for
// gnns/rgcn.py:84-114 (the reference trains through TF autodiff, models/sparse_graph_model.py:253-260).
//   d_agg = grad_out * act'(.) / div            (elementwise)
//   d_T[u, l, :] = sum_{(u->v) in A_l} s_{l,v} * d_agg[v, :]      (sorted-segment kernel over the reverse index)
//   d_H = d_T . [W_0|...|W_{L-1}]^T             (tcgen05 GEMM, transposed weight images)
//   d_W_l = H^T . d_T[:, l, :]                  (tiled FMA kernel, deterministic two-stage sum)
extern "C" int rgnn_rgcn_backward(const rgnn_plan_t* plan_c, const float* h, int32_t d_in, int32_t d_out,
                                  const float* const* edge_weights, const float* num_incoming, int activation,
                                  int aggregation, int normalize, const float* out, const float* grad_out,
                                  float* grad_h, float* const* grad_edge_weights, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  rgnn_plan* plan = const_cast<rgnn_plan*>(plan_c);   // the reverse index is built lazily inside the plan
  RGNN_REQUIRE(plan != nullptr && h != nullptr && out != nullptr && grad_out != nullptr && edge_weights != nullptr,
               "rgcn_backward: NULL argument");
  RGNN_REQUIRE(d_in > 0 && d_out > 0 && (d_in % 4) == 0 && (d_out % 4) == 0, "rgcn_backward: dims must be positive multiples of 4");
  RGNN_PROPAGATE(check_act(activation, "rgcn_backward"));
  RGNN_PROPAGATE(check_agg(aggregation, "rgcn_backward"));
  if (aggregation == RGNN_AGG_MAX) {
    set_error("rgcn_backward: the gradient of 'max' aggregation is not implemented in this build");
    return RGNN_E_UNSUPPORTED;
  }
  RGNN_REQUIRE(!normalize || num_incoming != nullptr, "rgcn_backward: normalize_by_num_incoming needs type_to_num_incoming_edges");
  const int V = plan->V, L = plan->L;
  RGNN_PROPAGATE(plan_ensure_reverse(plan, stream));
  Arena ar(workspace, workspace_bytes);
  float* d_agg = ar.floats((size_t)V * d_out);
  float* d_t = ar.floats((size_t)V * L * d_out);
  float* pre = nullptr;
  float* t_fwd = nullptr;
  if (activation == RGNN_ACT_GELU) {   // gelu' needs the pre-activation: recompute it (T = H.W, agg = segment reduce)
    pre = ar.floats((size_t)V * d_out);
    t_fwd = ar.floats((size_t)V * L * d_out);
  }
  static const bool gradw_fma = getenv("RGNN_GRADW_IMPL") != nullptr && strcmp(getenv("RGNN_GRADW_IMPL"), "fma") == 0;   // A/B reference
  float* gw_scratch = nullptr;
  if (grad_edge_weights) gw_scratch = ar.floats(gradw_fma ? grad_weight_scratch_floats(V, L, d_in, d_out) : gemm_tn_scratch_floats(d_in, L * d_out, V));
  RGNN_PROPAGATE(check_ws(ar, "rgcn_backward"));

  if (pre != nullptr) {
    RGNN_PROPAGATE(gemm_shared_a(ar, h, V, d_in, edge_weights, L, d_out, d_out, t_fwd, RGNN_ACT_LINEAR, stream));
    SegParams f;
    seg_from_plan(f, plan);
    f.D = d_out; f.table = t_fwd; f.stride_idx = (long)L * d_out; f.stride_type = d_out;
    f.num_incoming = normalize ? num_incoming : nullptr;
    f.agg = aggregation; f.out = pre; f.ld_out = d_out;
    RGNN_PROPAGATE(launch_seg_reduce(f, stream));
  }
  RGNN_PROPAGATE(launch_act_backward(grad_out, out, pre, V, d_out, activation, aggregation, plan->seg_off, d_agg, stream));
  {
    SegParams r;   // reverse index: segment = (source u, type l); gathered row = d_agg[original target]
    r.V = V * L; r.L = L; r.D = d_out;
    r.seg_off = plan->rev_seg_off; r.e_idx = plan->rev_src; r.e_type = plan->rev_type;
    r.table = d_agg; r.stride_idx = d_out; r.stride_type = 0;
    r.num_incoming = normalize ? num_incoming : nullptr; r.scale_ld = V; r.scale_by_idx = 1;
    r.heavy_list = plan->rev_heavy_list; r.heavy_count = plan->err_flag + 2;
    r.heavy_threshold = RGNN_HEAVY_SEGMENT; r.heavy_known = -1;
    r.agg = RGNN_AGG_SUM; r.out = d_t; r.ld_out = d_out;
    RGNN_PROPAGATE(launch_seg_reduce(r, stream));
  }
  if (grad_h != nullptr) {
    GemmParams g;
    g.A1 = d_t; g.lda1 = L * d_out; g.K1 = L * d_out;
    g.M = V; g.N = d_in; g.C = grad_h; g.ldc = d_in; g.ldb1 = d_out;
    g.batch_mode = BATCH_K_BLOCKS_T; g.batch = L; g.k_block = d_out;
    for (int l = 0; l < L; ++l) { g.bptr[l] = edge_weights[l]; g.bptr2[l] = nullptr; }
    RGNN_PROPAGATE(run_gemm(g, ar, stream));
  }
  if (grad_edge_weights != nullptr) {
    // dW_l = H^T . dT[:, l, :]  -- one TN contraction over the V nodes for all types (gemm_tn_tcgen05.cu)
    GradWTable tab;
    GemmTnOut tn;
    tn.block_cols = d_out; tn.ld = d_out;
    for (int l = 0; l < L; ++l) {
      RGNN_REQUIRE(grad_edge_weights[l] != nullptr && aligned16(grad_edge_weights[l]), "rgcn_backward: grad weight %d is NULL / misaligned", l);
      tab.out[l] = grad_edge_weights[l];
      tn.ptr[l] = grad_edge_weights[l];
    }
    if (gradw_fma) RGNN_PROPAGATE(launch_grad_weights(h, d_t, V, L, d_in, d_out, tab, gw_scratch, stream));
    else RGNN_PROPAGATE(launch_gemm_tn(h, d_in, d_t, L * d_out, d_in, L * d_out, V, tn, gw_scratch, stream));
  }
  return RGNN_OK;
}

// models/sparse_graph_model.py:176-191: for layer_idx in range(graph_num_layers): _apply_gnn_layer(...)
extern "C" int rgnn_rgcn_stack_forward(const rgnn_plan_t* plan, const float* h, int32_t d, int32_t num_layers,
                                       const float* const* edge_weights, const float* num_incoming, int activation,
                                       int aggregation, int normalize, float* out, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
  RGNN_REQUIRE(plan != nullptr && num_layers >= 1, "rgcn_stack: plan is NULL or num_layers < 1");
  RGNN_REQUIRE(edge_weights != nullptr, "rgcn_stack: edge_weights is NULL");
  const size_t row_bytes = align_up((size_t)plan->V * d * sizeof(float), 256);
  RGNN_REQUIRE(workspace != nullptr && workspace_bytes > 2 * row_bytes, "rgcn_stack: workspace too small");
  char* base = static_cast<char*>(workspace);
  float* buf[2] = {reinterpret_cast<float*>(base), reinterpret_cast<float*>(base + row_bytes)};
  void* inner = base + 2 * row_bytes;
  const size_t inner_bytes = workspace_bytes - 2 * row_bytes;
  const float* cur = h;
  for (int l = 0; l < num_layers; ++l) {
    float* dst = (l == num_layers - 1) ? out : buf[l & 1];
    RGNN_PROPAGATE(rgnn_rgcn_forward(plan, cur, d, d, edge_weights + (size_t)l * plan->L, num_incoming, activation,
                                     aggregation, normalize, 0, 1, dst, inner, inner_bytes, stream_));
    cur = dst;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// gnns/rgdcn.py:8-171
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_rgdcn_forward(const rgnn_plan_t* plan, const float* h, int32_t d, int32_t num_channels,
                                  const float* const* channel_weights, int use_full_state, const float* num_incoming,
                                  int activation, int aggregation, int normalize, int num_timesteps, float* out,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d, d, out, num_timesteps, "rgdcn"));
  RGNN_PROPAGATE(plan_wait_sources(plan, stream));
  RGNN_PROPAGATE(check_act(activation, "rgdcn"));
  RGNN_PROPAGATE(check_agg(aggregation, "rgdcn"));
  RGNN_REQUIRE(channel_weights != nullptr, "rgdcn: channel_weights is NULL");
  RGNN_REQUIRE(num_channels >= 1 && num_channels <= RGNN_MAX_EDGE_TYPES && (d % num_channels) == 0,
               "rgdcn: num_channels %d must divide the state dim %d (and be <= %d)", num_channels, d, RGNN_MAX_EDGE_TYPES);
  RGNN_REQUIRE(!normalize || num_incoming != nullptr, "rgdcn: normalize_by_num_incoming needs type_to_num_incoming_edges");
  const int V = plan->V, L = plan->L, C = num_channels, K = d / num_channels;
  RGNN_REQUIRE(K >= 4 && (K & (K - 1)) == 0 && K <= 128, "rgdcn: channel_dim %d must be a power of two in [4, 128]", K);
  for (int i = 0; i < L * C; ++i)
    RGNN_REQUIRE(channel_weights[i] != nullptr && aligned16(channel_weights[i]), "rgdcn: channel weight %d is NULL / misaligned", i);
  Arena ar(workspace, workspace_bytes);
  float* wdyn = ar.floats((size_t)V * L * d * K);
  float* buf[2] = {nullptr, nullptr};
  if (num_timesteps > 1) { buf[0] = ar.floats((size_t)V * d); buf[1] = ar.floats((size_t)V * d); }
  RGNN_PROPAGATE(check_ws(ar, "rgdcn"));
  const float* cur = h;
  for (int t = 0; t < num_timesteps; ++t) {
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    // W[v, l, c] = act(F_{l,c} . input_v) reshaped [K, K]  (:139-148; the Dense carries the layer's activation, :101-103)
    for (int l = 0; l < L; ++l) {
      GemmParams g;
      g.A1 = cur; g.lda1 = d; g.M = V; g.N = K * K; g.ldb1 = K * K;
      g.C = wdyn + (size_t)l * d * K; g.ldc = L * d * K;
      g.act = activation; g.batch = C;
      for (int c = 0; c < C; ++c) { g.bptr[c] = channel_weights[l * C + c]; g.bptr2[c] = nullptr; }
      if (use_full_state) { g.K1 = d; g.batch_mode = BATCH_SHARED_A; }       // input = the whole state h_v
      else { g.K1 = K; g.batch_mode = BATCH_COL_BLOCKS; }                    // input = the channel's slice h_v[c]
      RGNN_PROPAGATE(run_gemm(g, ar, stream));
    }
    RgdcnParams r;
    r.V = V; r.L = L; r.D = d; r.K = K;
    r.seg_off = plan->seg_off; r.e_src = plan->e_src; r.e_type = plan->e_type;
    r.h = cur; r.wdyn = wdyn; r.num_incoming = normalize ? num_incoming : nullptr;
    r.agg = aggregation; r.act_out = activation; r.out = dst;
    RGNN_PROPAGATE(launch_rgdcn_edges(r, stream));
    cur = dst;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// gnns/ggnn.py:8-95
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_ggnn_forward(const rgnn_plan_t* plan, const float* h, int32_t d_in, int32_t d_out,
                                 const float* const* edge_weights, const float* cell_kernel,
                                 const float* cell_recurrent_kernel, const float* cell_bias, int cell_kind,
                                 int activation, int aggregation, int num_timesteps, float* out, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d_in, d_out, out, num_timesteps, "ggnn"));
  RGNN_PROPAGATE(check_act(activation, "ggnn"));
  RGNN_PROPAGATE(check_agg(aggregation, "ggnn"));
  RGNN_REQUIRE(d_in == d_out, "ggnn: the recurrent cell needs state_dim == input dim (d_in=%d, d_out=%d)", d_in, d_out);
  if (cell_kind != RGNN_CELL_RNN && cell_kind != RGNN_CELL_GRU) {
    set_error("Unknown RNN cell type code %d.", cell_kind);                   // utils/utils.py:20
    return RGNN_E_INVALID;
  }
  RGNN_REQUIRE(edge_weights && cell_kernel && cell_recurrent_kernel && cell_bias, "ggnn: NULL weight pointer");
  RGNN_REQUIRE(aligned16(cell_kernel) && aligned16(cell_recurrent_kernel), "ggnn: cell kernels must be 16-byte aligned");
  const int V = plan->V, L = plan->L, D = d_out;
  for (int l = 0; l < L; ++l) RGNN_REQUIRE(edge_weights[l] != nullptr, "ggnn: edge weight %d is NULL", l);
  Arena ar(workspace, workspace_bytes);
  float* T = ar.floats((size_t)V * L * D);
  float* m = ar.floats((size_t)V * D);
  static const int slab_env = getenv("RGNN_GRU_SLAB") ? atoi(getenv("RGNN_GRU_SLAB")) : 0;   // rows per slab (experiment knob)
  const int slab_default = 148 * 128;                                         // one wave of 128-row tiles (measured: 18,944 rows 1.93 ms, 37,888 2.05, 56,832 2.23 on QM9-10k)
  const int slab = slab_env > 0 ? slab_env : (V < slab_default ? (V > 0 ? V : 1) : slab_default);
  float* z = ar.floats((size_t)slab * D);
  float* rh = ar.floats((size_t)slab * D);
  float* buf[2] = {ar.floats((size_t)V * D), ar.floats((size_t)V * D)};
  SegParams heavy;
  seg_heavy_scratch(heavy, plan, ar, D);
  RGNN_PROPAGATE(check_ws(ar, "ggnn"));

  const float* cur = h;
  for (int t = 0; t < num_timesteps; ++t) {                                   // ggnn.py:71
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    SegParams s;
    seg_from_plan(s, plan);
    s.D = D;
    RGNN_PROPAGATE(transform_sources(plan, ar, cur, D, D, edge_weights, T, stream, s));   // ggnn.py:80-82
    s.agg = aggregation; s.out = m; s.ld_out = D; s.heavy_scratch = heavy.heavy_scratch;   // ggnn.py:87-90
    // Experiment (RGNN_GRU_SLAB_EDGES=1; needs a batch without heavy targets): slab the edge stage together with the cell so
    // that a slab's aggregated messages are consumed out of L2.  Measured on QM9-10k: 1.975 ms vs 1.928 ms without (job L) --
    // five small edge-stage launches cost more than the saved round trip of m; off by default.
    static const bool slab_edges_env = getenv("RGNN_GRU_SLAB_EDGES") != nullptr && atoi(getenv("RGNN_GRU_SLAB_EDGES")) == 1;
    const bool slab_edges = slab_edges_env && cell_kind == RGNN_CELL_GRU && plan->num_heavy_host == 0;
    if (!slab_edges) RGNN_PROPAGATE(launch_seg_reduce(s, stream));
    GemmParams g;
    g.A1 = m; g.lda1 = D; g.K1 = D;
    g.M = plan->Vt; g.bias = cell_bias;   // the cell runs on the wanted target rows only
    if (cell_kind == RGNN_CELL_RNN) {                                         // SimpleRNNCell: act(x.W + b + h.U)
      g.A2 = cur; g.lda2 = D; g.K2 = D;
      g.B1 = cell_kernel; g.ldb1 = D; g.B2 = cell_recurrent_kernel; g.ldb2 = D;
      g.N = D; g.C = dst; g.ldc = D; g.epi = EPI_STORE; g.act = activation;
      RGNN_PROPAGATE(run_gemm(g, ar, stream));
    } else {                                                                  // GRUCell, gates z|r|h (A.4)
      // Two GEMMs per row SLAB: [z | r.h] = hs([m|h].[W_zr;U_zr] + b), then h' = z.h + (1-z).act([m | r.h].[W_h;U_h] + b_h).
      // z and r.h live in slab-sized scratch that every slab overwrites: with ~19k rows per slab (one wave of 128-row
      // tiles on 148 SMs) the slab's z, r.h, m, h and output rows (5 x 9.7 MB) stay in the 126 MB L2 between the two kernels and the
      // dirty z / r.h lines are overwritten before they are evicted -- the round trip through HBM of round 1
      // (4 x 92 MB per timestep on the QM9-10k batch) becomes L2 traffic.
      const int Vc = plan->Vt;
      for (int r0 = 0; r0 < Vc; r0 += slab) {
        const int rows = (Vc - r0 < slab) ? Vc - r0 : slab;
        const size_t off = (size_t)r0 * D;
        const float* m_slab = m + off;
        if (slab_edges) {                                  // aggregate this slab's targets into the first rows of m (reused by every slab)
          SegParams q = s;
          q.V = rows; q.seg_off = s.seg_off + r0; q.out = m;
          RGNN_PROPAGATE(launch_seg_reduce(q, stream));
          m_slab = m;
        }
        GemmParams g2 = g;
        g2.A1 = m_slab; g2.M = rows;
        g2.A2 = cur + off; g2.lda2 = D; g2.K2 = D;
        g2.B1 = cell_kernel; g2.ldb1 = 3 * D; g2.B2 = cell_recurrent_kernel; g2.ldb2 = 3 * D;
        g2.N = 2 * D; g2.C = z; g2.ldc = D; g2.C2 = rh; g2.ldc2 = D; g2.aux_h = cur + off; g2.ld_h = D;
        g2.epi = EPI_GRU_ZR;
        RGNN_PROPAGATE(run_gemm(g2, ar, stream));
        GemmParams o;
        o.A1 = m_slab; o.lda1 = D; o.K1 = D; o.A2 = rh; o.lda2 = D; o.K2 = D;
        o.B1 = cell_kernel + 2 * D; o.ldb1 = 3 * D; o.B2 = cell_recurrent_kernel + 2 * D; o.ldb2 = 3 * D;
        o.M = rows; o.N = D; o.bias = cell_bias + 2 * D; o.C = dst + off; o.ldc = D;
        o.aux_h = cur + off; o.ld_h = D; o.aux_z = z; o.ld_z = D;
        o.epi = EPI_GRU_OUT; o.act = activation;
        RGNN_PROPAGATE(run_gemm(o, ar, stream));
      }
    }
    cur = dst;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// gnns/rgat.py:9-141
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_rgat_forward(const rgnn_plan_t* plan, const float* h, int32_t d_in, int32_t d_out,
                                 const float* const* edge_weights, const float* const* attention, int num_heads,
                                 int activation, int num_timesteps, float* out, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d_in, d_out, out, num_timesteps, "rgat"));
  RGNN_PROPAGATE(plan_wait_sources(plan, stream));
  RGNN_PROPAGATE(check_act(activation, "rgat"));
  RGNN_REQUIRE(edge_weights != nullptr && attention != nullptr, "rgat: NULL weight table");
  RGNN_REQUIRE(num_heads >= 1 && (d_out % num_heads) == 0, "rgat: state_dim %d not divisible by num_heads %d", d_out, num_heads);
  const int V = plan->V, L = plan->L, D = d_out, K = num_heads;
  AttnTable at;
  for (int l = 0; l < L; ++l) {
    RGNN_REQUIRE(edge_weights[l] != nullptr && attention[l] != nullptr && aligned16(attention[l]), "rgat: weight %d is NULL / misaligned", l);
    at.att[l] = attention[l];
  }
  Arena ar(workspace, workspace_bytes);
  float* T = ar.floats((size_t)V * L * D);
  // per-edge logits: computed inside the edge kernel when a head's dh/4 lanes form a power-of-two group inside one warp
  const int dh = D / K, lph = dh / 4;
  const bool fused_scores = (dh % 4) == 0 && lph >= 1 && lph <= 32 && (lph & (lph - 1)) == 0 && getenv("RGNN_RGAT_UNFUSED") == nullptr;
  float* ssrc = fused_scores ? nullptr : ar.floats((size_t)V * L * K);
  float* stgt = fused_scores ? nullptr : ar.floats((size_t)V * L * K);
  float* buf[2] = {nullptr, nullptr};
  if (num_timesteps > 1) { buf[0] = ar.floats((size_t)V * D); buf[1] = ar.floats((size_t)V * D); }
  RGNN_PROPAGATE(check_ws(ar, "rgat"));

  const float* cur = h;
  int din = d_in;
  for (int t = 0; t < num_timesteps; ++t) {                                   // rgat.py:83
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    RGNN_PROPAGATE(gemm_shared_a(ar, cur, V, din, edge_weights, L, D, D, T, RGNN_ACT_LINEAR, stream));   // rgat.py:95-96
    if (!fused_scores) RGNN_PROPAGATE(launch_rgat_scores(T, V, L, D, K, at, ssrc, stgt, stream));   // rgat.py:106-115 (per node)
    RgatParams r;
    r.V = V; r.L = L; r.D = D; r.K = K; r.att = at;
    r.seg_off = plan->seg_off; r.e_src = plan->e_src; r.e_type = plan->e_type;
    r.table = T; r.s_src = ssrc; r.s_tgt = stgt; r.act_out = activation; r.out = dst;
    RGNN_PROPAGATE(launch_seg_rgat(r, stream));                                                      // rgat.py:120-138
    cur = dst; din = D;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// gnns/gnn_film.py:8-122
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_film_forward(const rgnn_plan_t* plan, const float* h, int32_t d_in, int32_t d_out,
                                 const float* const* edge_weights, const float* const* film_weights,
                                 const float* num_incoming, const float* ln_gamma, const float* ln_beta,
                                 int activation, int aggregation, int normalize, int num_timesteps, float* out,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d_in, d_out, out, num_timesteps, "gnn_film"));
  RGNN_PROPAGATE(check_act(activation, "gnn_film"));
  RGNN_PROPAGATE(check_agg(aggregation, "gnn_film"));
  RGNN_REQUIRE(edge_weights && film_weights && ln_gamma && ln_beta, "gnn_film: NULL weight pointer");
  RGNN_REQUIRE(aligned16(ln_gamma) && aligned16(ln_beta), "gnn_film: layer-norm parameters must be 16-byte aligned");
  RGNN_REQUIRE(!normalize || num_incoming != nullptr, "gnn_film: normalize_by_num_incoming needs type_to_num_incoming_edges");
  const int V = plan->V, L = plan->L, D = d_out;
  for (int l = 0; l < L; ++l) RGNN_REQUIRE(edge_weights[l] && film_weights[l], "gnn_film: weight %d is NULL", l);
  Arena ar(workspace, workspace_bytes);
  float* T = ar.floats((size_t)V * L * D);
  float* FW = ar.floats((size_t)V * L * 2 * D);
  float* buf[2] = {nullptr, nullptr};
  if (num_timesteps > 1) { buf[0] = ar.floats((size_t)V * D); buf[1] = ar.floats((size_t)V * D); }
  SegParams heavy;
  seg_heavy_scratch(heavy, plan, ar, D);
  RGNN_PROPAGATE(check_ws(ar, "gnn_film"));

  const float* cur = h;
  int din = d_in;
  for (int t = 0; t < num_timesteps; ++t) {                                   // gnn_film.py:85
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    // [gamma | beta] = F_l h_v for the wanted target rows (:102).  Experiment (RGNN_FILM_SLAB=1; plans without heavy targets):
    // row slabs small enough (<= 48 MB of gamma / beta rows) that the edge stage reads them back out of L2, target-side GEMM
    // and edge stage alternating per slab.  Measured on config 5 (50k / 1M): 0.441 ms vs 0.343 ms with ONE slab (job L) -- the
    // seven short GEMM / edge-stage launch pairs lose more to tails than L2 residency returns; off by default.
    const int Vc = plan->Vt;
    int slab = Vc > 0 ? Vc : 1;
    static const bool film_slab_env = getenv("RGNN_FILM_SLAB") != nullptr && atoi(getenv("RGNN_FILM_SLAB")) == 1;
    if (film_slab_env && plan->num_heavy_host == 0) {
      const long rows_fit = (48L << 20) / ((long)L * 2 * D * (long)sizeof(float));
      const long r = rows_fit / 128 * 128;
      if (r >= 1024 && r < slab) slab = (int)r;
    }
    // One slab (the default): the target-side GEMM goes FIRST -- it reads owned rows only, so on a sharded plan it overlaps a
    // pending halo exchange (rgnn_halo_exchange_overlapped), which transform_sources() below joins before touching halo rows.
    const bool fw_first = slab >= Vc;
    if (fw_first && Vc > 0)
      RGNN_PROPAGATE(gemm_shared_a(ar, cur, Vc, din, film_weights, L, 2 * D, 2 * D, FW, RGNN_ACT_LINEAR, stream));
    SegParams s;
    seg_from_plan(s, plan);
    s.D = D;
    RGNN_PROPAGATE(transform_sources(plan, ar, cur, din, D, edge_weights, T, stream, s));                        // :94 on nodes
    s.num_incoming = normalize ? num_incoming : nullptr;                      // :96-100
    s.msg_mode = MSG_FILM; s.mod_stride_node = (long)L * 2 * D; s.mod_stride_type = 2 * D;                        // :103-108
    s.act_msg = activation;                                                   // :112 (before the sum)
    s.agg = aggregation;                                                      // :113-116
    s.ln_gamma = ln_gamma + (size_t)t * D; s.ln_beta = ln_beta + (size_t)t * D;   // :120
    s.ld_out = D; s.heavy_scratch = heavy.heavy_scratch;
    for (int r0 = 0; r0 < Vc; r0 += slab) {
      const int rows = (Vc - r0 < slab) ? Vc - r0 : slab;
      if (!fw_first) RGNN_PROPAGATE(gemm_shared_a(ar, cur + (size_t)r0 * din, rows, din, film_weights, L, 2 * D, 2 * D, FW, RGNN_ACT_LINEAR, stream));
      SegParams q = s;
      q.V = rows; q.seg_off = s.seg_off + r0; q.mod_table = FW;             // FW holds this slab's rows from row 0
      if (q.num_incoming != nullptr) q.num_incoming = s.num_incoming + r0;    // c[type, v] = base[type * ld + v]
      q.out = dst + (size_t)r0 * D;
      RGNN_PROPAGATE(launch_seg_reduce(q, stream));
    }
    cur = dst; din = D;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// gnns/gnn_edge_mlp.py:7-122
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_edge_mlp_forward(const rgnn_plan_t* plan, const float* h, int32_t d_in, int32_t d_out,
                                     const float* const* mlp_kernels, const int32_t* mlp_dims,
                                     int num_edge_hidden_layers, const float* num_incoming, const float* ln_gamma,
                                     const float* ln_beta, int activation, int aggregation, int normalize,
                                     int use_target, int num_timesteps, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d_in, d_out, out, num_timesteps, "gnn_edge_mlp"));
  RGNN_PROPAGATE(plan_wait_sources(plan, stream));
  RGNN_PROPAGATE(check_act(activation, "gnn_edge_mlp"));
  RGNN_PROPAGATE(check_agg(aggregation, "gnn_edge_mlp"));
  RGNN_REQUIRE(mlp_kernels && mlp_dims && ln_gamma && ln_beta, "gnn_edge_mlp: NULL weight pointer");
  RGNN_REQUIRE(num_edge_hidden_layers >= 0, "gnn_edge_mlp: num_edge_hidden_layers %d < 0", num_edge_hidden_layers);
  RGNN_REQUIRE(!normalize || num_incoming != nullptr, "gnn_edge_mlp: normalize_by_num_incoming needs type_to_num_incoming_edges");
  const int nl = num_edge_hidden_layers + 1;
  RGNN_REQUIRE(nl <= RGNN_MAX_MLP_LAYERS && mlp_dims[nl] == d_out, "gnn_edge_mlp: MLP output dim %d != state_dim %d", mlp_dims[nl <= RGNN_MAX_MLP_LAYERS ? nl : 0], d_out);
  const int V = plan->V, D = d_out;
  Arena ar(workspace, workspace_bytes);
  float* buf[2] = {nullptr, nullptr};
  if (num_timesteps > 1) { buf[0] = ar.floats((size_t)V * D); buf[1] = ar.floats((size_t)V * D); }
  const size_t mark = ar.used;
  const float* cur = h;
  for (int t = 0; t < num_timesteps; ++t) {                                   // gnn_edge_mlp.py:84
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    ar.used = mark;                                                           // scratch of the previous timestep is dead
    MsgSource ms;
    RGNN_PROPAGATE(build_mlp_messages(plan, cur, d_in, mlp_kernels, mlp_dims, nl, use_target, RGNN_ACT_ELU, ar, stream, &ms));  // :76,:102
    RGNN_PROPAGATE(check_ws(ar, "gnn_edge_mlp"));
    SegParams s;
    seg_from_plan(s, plan);
    seg_from_source(s, ms);
    s.num_incoming = normalize ? num_incoming : nullptr;                      // :104-108
    s.act_msg = activation;                                                   // :112
    s.agg = aggregation;                                                      // :113-116
    s.ln_gamma = ln_gamma + (size_t)t * D; s.ln_beta = ln_beta + (size_t)t * D;   // :119
    s.out = dst; s.ld_out = D;
    RGNN_PROPAGATE(launch_seg_reduce(s, stream));
    cur = dst;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// gnns/rgin.py:7-142
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_rgin_forward(const rgnn_plan_t* plan, const float* h, int32_t d_in, int32_t d_out,
                                 const float* const* edge_mlp_kernels, const int32_t* edge_mlp_dims,
                                 int num_edge_mlp_hidden_layers, const float* const* aggr_kernels,
                                 const int32_t* aggr_dims, int num_aggr_mlp_hidden_layers, const float* ln_gamma,
                                 const float* ln_beta, int activation, int aggregation, int use_target,
                                 int num_timesteps, float* out, void* workspace, size_t workspace_bytes,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_PROPAGATE(check_common(plan, h, d_in, d_out, out, num_timesteps, "rgin"));
  RGNN_PROPAGATE(plan_wait_sources(plan, stream));
  RGNN_PROPAGATE(check_act(activation, "rgin"));
  RGNN_PROPAGATE(check_agg(aggregation, "rgin"));
  RGNN_REQUIRE(ln_gamma && ln_beta, "rgin: NULL layer-norm parameters");
  const int nl_edge = num_edge_mlp_hidden_layers < 0 ? 0 : num_edge_mlp_hidden_layers + 1;
  const int nl_aggr = num_aggr_mlp_hidden_layers < 0 ? 0 : num_aggr_mlp_hidden_layers + 1;
  RGNN_REQUIRE(nl_edge == 0 || (edge_mlp_kernels && edge_mlp_dims), "rgin: NULL edge MLP table");
  RGNN_REQUIRE(nl_aggr == 0 || (aggr_kernels && aggr_dims), "rgin: NULL aggregation MLP table");
  RGNN_REQUIRE(nl_aggr <= RGNN_MAX_MLP_LAYERS, "rgin: aggregation MLP too deep");
  const int V = plan->V, D = d_out;
  Arena ar(workspace, workspace_bytes);
  float* buf[2] = {nullptr, nullptr};
  if (num_timesteps > 1) { buf[0] = ar.floats((size_t)V * D); buf[1] = ar.floats((size_t)V * D); }
  const size_t mark = ar.used;
  const float* cur = h;
  for (int t = 0; t < num_timesteps; ++t) {                                   // rgin.py:103
    float* dst = (t == num_timesteps - 1) ? out : buf[t & 1];
    ar.used = mark;
    MsgSource ms;
    RGNN_PROPAGATE(build_mlp_messages(plan, cur, d_in, edge_mlp_kernels, edge_mlp_dims, nl_edge, use_target, activation, ar, stream, &ms));  // :95,:122
    RGNN_PROPAGATE(check_ws(ar, "rgin"));
    const int width = ms.width;
    SegParams s;
    seg_from_plan(s, plan);
    seg_from_source(s, ms);
    s.act_msg = (nl_edge > 0) ? activation : RGNN_ACT_LINEAR;                 // :128-129
    s.agg = aggregation;                                                      // :130-133
    if (nl_aggr == 0) {
      RGNN_REQUIRE(width == D, "rgin: message width %d != state_dim %d and no aggregation MLP maps it", width, D);
      s.act_out = activation;                                                 // :138
      s.ln_gamma = ln_gamma + (size_t)t * D; s.ln_beta = ln_beta + (size_t)t * D;   // :139
      s.out = dst; s.ld_out = D;
      RGNN_PROPAGATE(launch_seg_reduce(s, stream));
    } else {
      RGNN_REQUIRE(aggr_dims[0] == width && aggr_dims[nl_aggr] == D, "rgin: aggregation MLP dims [%d .. %d] do not match [%d .. %d]",
                   aggr_dims[0], aggr_dims[nl_aggr], width, D);
      float* agg = ar.floats((size_t)V * width);
      RGNN_PROPAGATE(check_ws(ar, "rgin"));
      s.out = agg; s.ld_out = width;
      RGNN_PROPAGATE(launch_seg_reduce(s, stream));
      const float* prev = agg;
      for (int j = 0; j < nl_aggr; ++j) {                                     // :136-137 (+ :138 fused into the last layer)
        RGNN_REQUIRE(aggr_kernels[j] != nullptr && (aggr_dims[j + 1] % 4) == 0, "rgin: aggregation MLP layer %d invalid", j);
        float* next = ar.floats((size_t)V * aggr_dims[j + 1]);
        RGNN_PROPAGATE(check_ws(ar, "rgin"));
        GemmParams g;
        g.A1 = prev; g.lda1 = aggr_dims[j]; g.K1 = aggr_dims[j];
        g.B1 = aggr_kernels[j]; g.ldb1 = aggr_dims[j + 1];
        g.M = V; g.N = aggr_dims[j + 1]; g.C = next; g.ldc = aggr_dims[j + 1];
        g.act = activation;   // hidden layers: MLP activation (rgin.py:80); last layer: the explicit activation of :138
        RGNN_PROPAGATE(run_gemm(g, ar, stream));
        prev = next;
      }
      RGNN_PROPAGATE(launch_layer_norm(prev, V, D, ln_gamma + (size_t)t * D, ln_beta + (size_t)t * D, dst, stream));  // :139
    }
    cur = dst;
  }
  return RGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// building blocks
// ---------------------------------------------------------------------------------------------
extern "C" int rgnn_segment_aggregate(const rgnn_plan_t* plan, const float* data, int32_t d, int aggregation,
                                      float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_REQUIRE(plan != nullptr && data != nullptr && out != nullptr, "segment_aggregate: NULL argument");
  RGNN_PROPAGATE(check_agg(aggregation, "segment_aggregate"));
  SegParams s;
  seg_from_plan(s, plan);
  s.e_idx = plan->e_orig; s.table = data; s.stride_idx = d; s.stride_type = 0; s.D = d;
  s.agg = aggregation; s.out = out; s.ld_out = d;
  return launch_seg_reduce(s, stream);
}

// The edge stage on per-node transformed states (rgcn.py:84-112, ggnn.py:76-90 after re-association):
// out[v] = agg_{l, (u,v) in A_l} s_{l,v} * table[u, l, :]
extern "C" int rgnn_edge_aggregate_forward(const rgnn_plan_t* plan, const float* table, int32_t d, const float* num_incoming,
                                           int aggregation, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_REQUIRE(plan != nullptr && table != nullptr && out != nullptr, "edge_aggregate: NULL argument");
  RGNN_REQUIRE(d > 0 && (d % 4) == 0 && aligned16(table) && aligned16(out), "edge_aggregate: d must be a positive multiple of 4, rows 16-byte aligned");
  RGNN_PROPAGATE(check_agg(aggregation, "edge_aggregate"));
  SegParams s;
  seg_from_plan(s, plan);
  s.D = d; s.table = table; s.stride_idx = (long)plan->L * d; s.stride_type = d;
  s.num_incoming = num_incoming;
  s.agg = aggregation; s.out = out; s.ld_out = d;
  return launch_seg_reduce(s, stream);
}

// d_table[u, l, :] = sum_{(u,v) in A_l} s_{l,v} * grad_out[v, :] / div(v)   (reverse index: segments = (source, type))
extern "C" int rgnn_edge_aggregate_backward(const rgnn_plan_t* plan_c, const float* grad_out, int32_t d,
                                            const float* num_incoming, int aggregation, float* d_table, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  rgnn_plan* plan = const_cast<rgnn_plan*>(plan_c);
  RGNN_REQUIRE(plan != nullptr && grad_out != nullptr && d_table != nullptr, "edge_aggregate_backward: NULL argument");
  RGNN_REQUIRE(d > 0 && (d % 4) == 0 && aligned16(grad_out) && aligned16(d_table), "edge_aggregate_backward: d must be a positive multiple of 4, rows 16-byte aligned");
  RGNN_PROPAGATE(check_agg(aggregation, "edge_aggregate_backward"));
  if (aggregation == RGNN_AGG_MAX) {
    set_error("edge_aggregate_backward: the gradient of 'max' aggregation is not implemented in this kernel");
    return RGNN_E_UNSUPPORTED;
  }
  RGNN_PROPAGATE(plan_ensure_reverse(plan, stream));
  const int V = plan->V, L = plan->L;
  const float* d_agg = grad_out;
  float* scratch = nullptr;
  if (aggregation != RGNN_AGG_SUM) {   // mean / sqrt_n: divide by the segment size first
    RGNN_CHECK_CUDA(cudaMallocAsync(&scratch, sizeof(float) * (size_t)(V > 0 ? V : 1) * d, stream));
    const int rc = launch_act_backward(grad_out, grad_out, nullptr, V, d, RGNN_ACT_LINEAR, aggregation, plan->seg_off, scratch, stream);
    if (rc != RGNN_OK) { cudaFreeAsync(scratch, stream); return rc; }
    d_agg = scratch;
  }
  SegParams r;
  r.V = V * L; r.L = L; r.D = d;
  r.seg_off = plan->rev_seg_off; r.e_idx = plan->rev_src; r.e_type = plan->rev_type;
  r.table = d_agg; r.stride_idx = d; r.stride_type = 0;
  r.num_incoming = num_incoming; r.scale_ld = V; r.scale_by_idx = 1;
  r.heavy_list = plan->rev_heavy_list; r.heavy_count = plan->err_flag + 2;
  r.heavy_threshold = RGNN_HEAVY_SEGMENT; r.heavy_known = -1;
  r.agg = RGNN_AGG_SUM; r.out = d_table; r.ld_out = d;
  const int rc = launch_seg_reduce(r, stream);
  if (scratch != nullptr) cudaFreeAsync(scratch, stream);
  return rc;
}

// Scratch of rgnn_dense_forward / rgnn_dense_backward: the weight images of the contraction(s) + the split-K partial tiles.
extern "C" size_t rgnn_dense_workspace_bytes(int32_t m, int32_t k, int32_t n) {
  if (m < 0 || k <= 0 || n <= 0) return 0;
  GemmParams f;
  f.M = m; f.N = n; f.K1 = k; f.lda1 = k; f.ldb1 = n; f.ldc = n;
  GemmParams t;
  t.M = m; t.N = k; t.K1 = n; t.lda1 = n; t.ldb1 = n; t.ldc = k; t.batch_mode = BATCH_K_BLOCKS_T; t.batch = 1; t.k_block = n;
  const size_t pack = gemm_tc_pack_bytes_uncached(f) > gemm_tc_pack_bytes_uncached(t) ? gemm_tc_pack_bytes_uncached(f) : gemm_tc_pack_bytes_uncached(t);
  return align_up(pack, 256) + align_up(gemm_tn_scratch_floats(k, n, m) * sizeof(float), 256) + 512;
}

extern "C" int rgnn_dense_forward(const float* a, int32_t m, int32_t k, const float* b, int32_t n, const float* bias,
                                  int activation, float* c, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_REQUIRE(a && b && c, "dense: NULL argument");
  RGNN_PROPAGATE(check_act(activation, "dense"));
  GemmParams g;
  g.A1 = a; g.lda1 = k; g.K1 = k; g.B1 = b; g.ldb1 = n; g.M = m; g.N = n; g.C = c; g.ldc = n;
  g.bias = bias; g.act = activation;
  Arena ar(workspace, workspace_bytes);
  return run_gemm(g, ar, stream);
}

// gradients of the linear map C = A . B of rgnn_dense_forward (TF autodiff of tf.keras Dense, sparse_graph_model.py:253)
extern "C" int rgnn_dense_backward(const float* a, int32_t m, int32_t k, const float* b, int32_t n, const float* grad_c,
                                   float* grad_a, float* grad_b, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  RGNN_REQUIRE((grad_c != nullptr || m == 0) && m >= 0 && k > 0 && n > 0 && (k % 4) == 0 && (n % 4) == 0, "dense_backward: bad arguments (m=%d k=%d n=%d)", m, k, n);
  Arena ar(workspace, workspace_bytes);
  if (grad_a != nullptr && m > 0) {   // dA = dC . B^T
    RGNN_REQUIRE(b != nullptr, "dense_backward: grad_a needs b");
    GemmParams g;
    g.A1 = grad_c; g.lda1 = n; g.K1 = n; g.M = m; g.N = k; g.C = grad_a; g.ldc = k; g.ldb1 = n;
    g.batch_mode = BATCH_K_BLOCKS_T; g.batch = 1; g.k_block = n; g.bptr[0] = b; g.bptr2[0] = nullptr;
    RGNN_PROPAGATE(run_gemm(g, ar, stream));
  }
  if (grad_b != nullptr) {            // dB = A^T . dC
    RGNN_REQUIRE(a != nullptr || m == 0, "dense_backward: grad_b needs a");
    GemmTnOut tn;
    tn.block_cols = n; tn.ld = n; tn.ptr[0] = grad_b;
    float* ws = ar.floats(gemm_tn_scratch_floats(k, n, m));
    RGNN_PROPAGATE(check_ws(ar, "dense_backward"));
    RGNN_PROPAGATE(launch_gemm_tn(a, k, grad_c, n, k, n, m, tn, ws, stream));
  }
  return RGNN_OK;
}

extern "C" int rgnn_layer_norm(const float* x, int32_t rows, int32_t d, const float* gamma, const float* beta,
                               float* out, void* stream_) {
  RGNN_REQUIRE(x && gamma && beta && out, "layer_norm: NULL argument");
  return launch_layer_norm(x, rows, d, gamma, beta, out, static_cast<cudaStream_t>(stream_));
}
