// seg.cuh -- edge-stage kernels: gather message rows by source, modulate, segment-reduce to the target.
#pragma once
#include "common.cuh"
#include "plan.cuh"

namespace rgnn {

enum MsgMode {
  MSG_LINEAR = 0,   // m = s * t                      (RGCN / GGNN / RGIN / materialised per-edge MLP outputs)
  MSG_FILM = 1,     // m = gamma * (s * t) + beta     (gnn_film.py:96-108)
  MSG_ADDTGT = 2,   // m = s * (t + q)                ([h_u | h_v] . W  ==  h_u.W_src + h_v.W_tgt : rgcn.py:91-96, gnn_edge_mlp.py:95-102)
};

struct SegParams {
  int V = 0, L = 1, D = 0;
  const int32_t* seg_off = nullptr;
  const int32_t* e_idx = nullptr;      // plan->e_src (tables indexed by node) or plan->e_orig (per-edge matrices)
  const int32_t* e_type = nullptr;
  const float* table = nullptr;        // message row of edge e = table + e_idx[e]*stride_idx + e_type[e]*stride_type
  long stride_idx = 0, stride_type = 0;
  const float* num_incoming = nullptr; // [L, scale_ld] fp32 -> s = 1/(c + 1e-7) (rgcn.py:100-104); NULL -> s = 1
  int scale_ld = 0;                    // row length of num_incoming (number of graph nodes)
  int scale_by_idx = 0;                // 0: c[type, v] of the segment's target v; 1: c[type, e_idx] (backward: the edge's original target)
  int msg_mode = MSG_LINEAR;
  const float* mod_table = nullptr;    // FILM: gamma at +0, beta at +D ; ADDTGT: q.  row = mod_table + v*mod_stride_node + type*mod_stride_type
  long mod_stride_node = 0, mod_stride_type = 0;
  int act_msg = RGNN_ACT_LINEAR;       // activation applied per message before the reduction
  int agg = RGNN_AGG_SUM;
  int act_out = RGNN_ACT_LINEAR;       // activation applied to the aggregate
  const float* ln_gamma = nullptr;     // non-NULL -> tf.contrib.layers.layer_norm epilogue (eps 1e-12)
  const float* ln_beta = nullptr;
  float* out = nullptr;
  int ld_out = 0;
  // degree-skew handling: targets with more than heavy_threshold incoming edges are skipped by the warp-per-target
  // kernel and reduced by a whole CTA each (seg_reduce_heavy_kernel).  heavy_list / heavy_count live in the plan.
  const int32_t* heavy_list = nullptr;
  const int* heavy_count = nullptr;    // device counter
  int heavy_threshold = 0;             // 0 = no splitting
  int heavy_known = -1;                // host copy of *heavy_count, or -1 when it was never read back
  // multi-CTA split (forward plans; needs scratch): every RGNN_HEAVY_CHUNK edges of a heavy target are one work item reduced
  // by one CTA into heavy_scratch[item, :]; a second kernel adds a target's partial rows in a fixed order and finishes the row
  const int32_t* heavy_base = nullptr; // [heavy targets] first item of heavy target i
  const int32_t* heavy_items = nullptr;// (target, chunk) pairs
  const int* heavy_item_count = nullptr;
  int heavy_items_known = -1, heavy_items_cap = 0, heavy_chunk = 0;
  float* heavy_scratch = nullptr;      // [heavy_items_cap, D] floats; NULL -> one CTA per heavy target (no split)
};
int launch_seg_reduce(const SegParams& p, cudaStream_t stream);

struct AttnTable { const float* att[RGNN_MAX_EDGE_TYPES]; };
struct RgatParams {
  int V = 0, L = 1, D = 0, K = 1;
  AttnTable att;                       // per-type attention vectors [2D] (fused-score path: s_src / s_tgt are NULL)
  const int32_t* seg_off = nullptr;
  const int32_t* e_src = nullptr;
  const int32_t* e_type = nullptr;
  const float* table = nullptr;        // T [V, L, D]
  const float* s_src = nullptr;        // [V, L, K]
  const float* s_tgt = nullptr;        // [V, L, K]
  int act_out = RGNN_ACT_LINEAR;
  float* out = nullptr;
};
int launch_seg_rgat(const RgatParams& p, cudaStream_t stream);

// gnns/rgdcn.py:121-171: messages h_u[c,:] . W[v,l,c] with a per-(target, type, channel) K x K kernel computed from the
// target's state.  wdyn [V, L, C, K, K] row-major (C = D / K).
struct RgdcnParams {
  int V = 0, L = 1, D = 0, K = 0;
  const int32_t* seg_off = nullptr;
  const int32_t* e_src = nullptr;
  const int32_t* e_type = nullptr;
  const float* h = nullptr;            // [V, D]
  const float* wdyn = nullptr;         // [V, L, D * K]
  const float* num_incoming = nullptr; // [L, V] or NULL
  int agg = RGNN_AGG_SUM;
  int act_out = RGNN_ACT_LINEAR;
  float* out = nullptr;                // [V, D]
};
int launch_rgdcn_edges(const RgdcnParams& p, cudaStream_t stream);

// s_src[n,l,k] = <att_l[k*2d : k*2d+d], T[n,l,k*d:(k+1)*d]>, s_tgt with att_l[k*2d+d : (k+1)*2d]  (rgat.py:106-115)
int launch_rgat_scores(const float* table, int V, int L, int D, int K, const AttnTable& att, float* s_src,
                       float* s_tgt, cudaStream_t stream);

// X[i, :] (i in type-major original order) = act(P[src_i, type, :] + Q[tgt_i, type, :])   (pq != NULL), or
// X[i, :] = [h[src_i] | h[tgt_i]]                                                         (concat mode)
struct EdgeBuildParams {
  int L = 1, D = 0;                     // D = width of P / Q rows (or of h in concat mode)
  const int32_t* o_src = nullptr;
  const int32_t* o_tgt = nullptr;
  int32_t type_off[RGNN_MAX_EDGE_TYPES + 1];
  int max_type_edges = 0;
  const float* p = nullptr; long p_stride_node = 0, p_stride_type = 0;
  const float* q = nullptr; long q_stride_node = 0, q_stride_type = 0;   // NULL -> concat mode
  int concat = 0;
  int act = RGNN_ACT_LINEAR;
  float* x = nullptr; int ldx = 0;
};
int launch_edge_build(const EdgeBuildParams& p, cudaStream_t stream);

int launch_layer_norm(const float* x, int rows, int D, const float* gamma, const float* beta, float* out,
                      cudaStream_t stream);

// d_agg[v, :] = grad_out[v, :] * act'(.) / div(v)   -- the elementwise head of every layer backward.
// act' is evaluated from the forward OUTPUT for linear/tanh/relu/leaky_relu/elu/selu and from the
// pre-activation `pre` (must be non-NULL) for gelu.  div: 1 (sum/max), n (mean), sqrt(n) (sqrt_n), n = max(in-degree, 1).
int launch_act_backward(const float* grad_out, const float* out, const float* pre, int V, int D, int act, int agg,
                        const int32_t* seg_off, float* d_agg, cudaStream_t stream);

// grad_w[l][i][j] = sum_v h[v, i] * d_t[v, l, j]   (d_t is [V, L, D], h is [V, Din]); deterministic two-stage sum.
struct GradWTable { float* out[RGNN_MAX_EDGE_TYPES]; };
size_t grad_weight_scratch_floats(int V, int L, int d_in, int d_out);
int launch_grad_weights(const float* h, const float* d_t, int V, int L, int d_in, int d_out, const GradWTable& out,
                        float* scratch, cudaStream_t stream);

}  // namespace rgnn
