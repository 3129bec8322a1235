// plan.cuh -- device-resident graph structure of one batch (see rgnn_plan_create in include/rgnn.h).
#pragma once
#include "common.cuh"

struct rgnn_plan {
  int32_t V = 0;
  int32_t Vt = 0;               // rows [0, Vt) are the targets whose outputs are wanted (rgnn_plan_set_num_targets); default V
  int32_t L = 0;
  int64_t M = 0;
  // CSR by target over all edge types; incoming edges of v sorted by (type, original position)
  int32_t* seg_off = nullptr;   // [V+1]
  int32_t* e_src = nullptr;     // [M]
  int32_t* e_type = nullptr;    // [M]
  int32_t* e_orig = nullptr;    // [M] position in the type-major concatenation of the inputs
  // the inputs, concatenated type-major (rows of per-edge matrices live in this order)
  int32_t* o_src = nullptr;     // [M]
  int32_t* o_tgt = nullptr;     // [M]
  int32_t type_off[RGNN_MAX_EDGE_TYPES + 1] = {0};   // host copy: block of type l = [type_off[l], type_off[l+1])
  int32_t max_type_edges = 0;
  int device = 0;
  // reverse index for the backward pass (built lazily by plan_ensure_reverse): CSR over (source, type) pairs,
  // segment id = src * L + type; rev_src[e] = the edge's ORIGINAL target, rev_type[e] = its type
  int32_t* rev_seg_off = nullptr;   // [V*L + 1]
  int32_t* rev_src = nullptr;       // [M]
  int32_t* rev_type = nullptr;      // [M]
  void* rev_block = nullptr;
  // Compact transform table (sparsely typed graphs): the distinct (source, type) pairs that occur as edges, ordered by
  // (type, source).  Built at plan creation when M < 0.75 * V * L (then at most that share of the V*L rows of
  // T = H.[W_0|..|W_{L-1}] is ever gathered): the per-type Dense is applied to these rows only (row-range GEMM whose A rows
  // follow pair_src) and the edge stage addresses the compact table through e_pair.  QM9-10k: 1.10 of 4 rows per node.
  int32_t* pair_src = nullptr;      // [n_pairs] source node of compact row r
  int32_t* e_pair = nullptr;        // [M] compact row of every edge, CSR-by-target order (parallel to e_src / e_type)
  int32_t* pair_off_dev = nullptr;  // [L+1] device copy of pair_type_off
  int32_t pair_type_off[RGNN_MAX_EDGE_TYPES + 1] = {0};   // host: rows of type l = [pair_type_off[l], pair_type_off[l+1])
  int32_t n_pairs = -1;             // host: -1 = no pair table / offsets not read back yet (deferred check)
  int32_t max_type_pairs = 0;
  void* pair_block = nullptr;
  // targets with more than RGNN_HEAVY_SEGMENT incoming edges (reduced by a whole CTA, see seg_kernels.cu)
  int32_t* heavy_list = nullptr;    // [V]
  int32_t* rev_heavy_list = nullptr; // [V*L] (reverse index)
  // multi-CTA split of the heavy targets: target i of heavy_list owns the work items [heavy_base[i], heavy_base[i] + n_i),
  // item j = the j-th RGNN_HEAVY_CHUNK edges of the target's segment; heavy_items[k] = (target, chunk) of item k
  int32_t* heavy_base = nullptr;    // [V] first item of heavy target i
  int32_t* heavy_items = nullptr;   // [2 * heavy_items_cap] (target, chunk) pairs
  int32_t heavy_items_cap = 0;      // upper bound of the item count (3 M / 512 + 2)
  int num_heavy_items_host = -1;    // host copy of flags[3] once rgnn_plan_status has read it
  int num_heavy_host = -1;          // host copy of flags[1] once rgnn_plan_status has read it, else -1
  int* err_flag = nullptr;          // device flags: [0] out-of-range node id seen, [1] number of heavy targets, [2] heavy (source,type) pairs      // device flag: an adjacency list held an out-of-range node id
  // Set by rgnn_halo_exchange_overlapped: rows >= Vt (the halo) become valid when this event fires.  The next layer forward
  // makes its stream wait for it right before the first kernel that reads halo rows (after the target-side work) and clears it.
  mutable cudaEvent_t source_ready = nullptr;
  void* block = nullptr;        // the one pool allocation behind all arrays above
  cudaStream_t stream = nullptr; // creation stream (the block is freed stream-ordered on it)
};

namespace rgnn {
// Build plan->rev_* on `stream` if absent (not thread-safe; called by the first backward on this plan).
int plan_ensure_reverse(rgnn_plan* plan, cudaStream_t stream);
// make `stream` wait for a pending overlapped halo exchange (no-op otherwise)
inline int plan_wait_sources(const rgnn_plan* plan, cudaStream_t stream) {
  if (plan->source_ready != nullptr) {
    const cudaError_t e = cudaStreamWaitEvent(stream, plan->source_ready, 0);
    plan->source_ready = nullptr;
    if (e != cudaSuccess) { set_error("CUDA error %s waiting for the halo exchange: %s", cudaGetErrorName(e), cudaGetErrorString(e)); return RGNN_E_CUDA; }
  }
  return RGNN_OK;
}
constexpr int RGNN_HEAVY_SEGMENT = 512;
constexpr int RGNN_HEAVY_CHUNK = 256;     // edges per work item of a split heavy target (one CTA each)
}  // namespace rgnn
