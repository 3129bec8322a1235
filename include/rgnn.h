/*
 * rgnn.h -- C ABI of the B200-native relational message-passing engine (librgnn.so).
 *
 * The reference (microsoft/tf-gnn-samples) has no FFI: its boundary is the Python call
 * convention  Sparse_Graph_Model._apply_gnn_layer(node_representations, adjacency_lists,
 * type_to_num_incoming_edges, num_timesteps)  (models/sparse_graph_model.py:204-225), served by
 * gnns.sparse_<x>_layer(...) (gnns/__init__.py:1-7).  Each rgnn_<x>_forward below is what a
 * ctypes binding of that layer function calls; the argument lists mirror the keyword
 * arguments the model adapters pass (models/<x>_model.py), plus explicit weight pointers
 * (the reference creates its weights inside the layer function under a tf.variable_scope).
 *
 * Conventions
 *   return    0 = OK, negative = error (RGNN_E_*); message via rgnn_last_error() (thread-local).
 *             Nothing throws across the ABI, nothing calls exit().
 *   memory    every data pointer is a DEVICE pointer owned by the caller (16-byte aligned,
 *             row-major fp32 / int32); "host array" arguments are small host-side tables
 *             (pointer lists, sizes).  The library borrows pointers for the duration of the
 *             call only.  Plans own their index buffers.
 *   async     all work is enqueued on `stream` (a cudaStream_t passed as void*); forwards
 *             make no hidden synchronisation and are CUDA-Graph capturable.  plan_create
 *             synchronises the stream once (it sizes its buffers on the host).
 *   threads   re-entrant on distinct plans/streams; one plan must not be used concurrently.
 *   layouts   node states [V, D] fp32 row-major, D % 4 == 0; adjacency lists int32 [E_l, 2]
 *             (col 0 = source, col 1 = target: gnns/rgcn.py:85-86); in-degrees fp32 [L, V]
 *             (tasks/sparse_graph_task.py:145); weights in Keras orientation kernel[in, out]
 *             (y = x . kernel) so reference checkpoints load without transposition.
 */
#ifndef RGNN_H_
#define RGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RGNN_API __attribute__((visibility("default")))
#else
#define RGNN_API
#endif

#define RGNN_VERSION 200          /* 0.2.0 */
#define RGNN_MAX_EDGE_TYPES 64
#define RGNN_MAX_MLP_LAYERS 8
#define RGNN_MAX_STATE_DIM 512    /* per-row register tile of the segment kernels */
#define RGNN_MAX_WORLD 16         /* ranks of one node-range partition (one NVSwitch domain) */
#define RGNN_PEER_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

/* error codes */
#define RGNN_OK 0
#define RGNN_E_INVALID (-1)       /* bad argument (NULL, misaligned, dim % 4 != 0, unknown enum ...) */
#define RGNN_E_CUDA (-2)          /* CUDA runtime error; message carries cudaGetErrorString */
#define RGNN_E_WORKSPACE (-3)     /* workspace too small */
#define RGNN_E_UNSUPPORTED (-4)   /* valid in the reference but outside this build's limits */

/* utils/utils.py:36-58 get_activation (names lower-cased there) */
enum rgnn_activation {
  RGNN_ACT_LINEAR = 0, RGNN_ACT_TANH = 1, RGNN_ACT_RELU = 2, RGNN_ACT_LEAKY_RELU = 3,
  RGNN_ACT_ELU = 4, RGNN_ACT_SELU = 5, RGNN_ACT_GELU = 6
};
/* utils/utils.py:23-33 get_aggregation_function */
enum rgnn_aggregation { RGNN_AGG_SUM = 0, RGNN_AGG_MAX = 1, RGNN_AGG_MEAN = 2, RGNN_AGG_SQRT_N = 3 };
/* utils/utils.py:10-20 get_gated_unit (LSTM is unusable as the reference calls it: ggnn.py:92) */
enum rgnn_cell { RGNN_CELL_RNN = 0, RGNN_CELL_GRU = 1 };
enum rgnn_layer_kind {
  RGNN_LAYER_RGCN = 0, RGNN_LAYER_GGNN = 1, RGNN_LAYER_RGAT = 2, RGNN_LAYER_FILM = 3,
  RGNN_LAYER_EDGE_MLP = 4, RGNN_LAYER_RGIN = 5, RGNN_LAYER_RGCN_BACKWARD = 6,
  RGNN_LAYER_RGDCN = 7   /* rgnn_workspace_bytes: pass channel_dim as mlp_layers */
};

typedef struct rgnn_plan rgnn_plan_t;
typedef struct rgnn_halo_plan rgnn_halo_plan_t;

RGNN_API int rgnn_version(void);
RGNN_API const char* rgnn_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
RGNN_API int64_t rgnn_launch_count(void);

/*
 * Plan = the batch's graph structure in the layout the kernels want, built once per batch and
 * reused by every layer and timestep (replaces the per-layer tf.concat of targets + unsorted
 * segment ids: gnns/rgcn.py:76-78,108-112).  Contents (device): CSR by target over ALL edge
 * types, incoming edges of a node sorted by (type, original position) -- a stable sort, so
 * every reduction order is deterministic.
 *   adjacency_lists : host array of L device pointers, each int32 [E_l, 2]
 *   num_edges       : host array [L] (E_l may be 0: tasks/ppi_task.py:246-249)
 */
RGNN_API int rgnn_plan_create(rgnn_plan_t** out, int32_t num_nodes, int32_t num_edge_types,
                     const int32_t* const* adjacency_lists, const int64_t* num_edges, void* stream);
/* Same, with flags.  RGNN_PLAN_DEFERRED_CHECK: do not synchronise; the index-range check result stays on the
 * device until rgnn_plan_status() (which synchronises the creation stream) is called.  Out-of-range ids are
 * clamped to node 0 inside the plan, so later kernels stay memory-safe either way. */
#define RGNN_PLAN_DEFERRED_CHECK 1
RGNN_API int rgnn_plan_create_ex(rgnn_plan_t** out, int32_t num_nodes, int32_t num_edge_types,
                        const int32_t* const* adjacency_lists, const int64_t* num_edges, int flags, void* stream);
RGNN_API int rgnn_plan_status(const rgnn_plan_t* plan);
/* Sharded execution (one rank of a node-range partition: owned nodes first, halo nodes after them): declare that only
 * rows [0, num_targets) are targets whose outputs are wanted.  The edge stage then reduces only those rows and the
 * TARGET-side node-level work of the layers (FiLM's gamma/beta GEMM, GGNN's cell) runs on them only; source-side transforms
 * still cover all V rows.  Output rows >= num_targets are left untouched; layers must be called with num_timesteps == 1
 * (halo states are refreshed by the caller's exchange between steps).  Default: num_targets = V. */
RGNN_API int rgnn_plan_set_num_targets(rgnn_plan_t* plan, int32_t num_targets);
RGNN_API int rgnn_plan_destroy(rgnn_plan_t* plan);
RGNN_API int32_t rgnn_plan_num_nodes(const rgnn_plan_t* plan);
RGNN_API int32_t rgnn_plan_num_edge_types(const rgnn_plan_t* plan);
RGNN_API int64_t rgnn_plan_num_edges(const rgnn_plan_t* plan);       /* M = sum_l E_l */
/* Copy the plan's arrays into caller DEVICE buffers (any may be NULL): seg_off [V+1],
 * e_src [M], e_type [M], e_orig [M] (position in the type-major concatenation of the inputs). */
RGNN_API int rgnn_plan_export(const rgnn_plan_t* plan, int32_t* seg_off, int32_t* e_src, int32_t* e_type,
                     int32_t* e_orig, void* stream);

/* Static-weight mode (inference / benchmarking), off by default.  The tensor-core GEMM consumes the
 * weights as pre-swizzled TF32 hi/lo shared-memory images; normally they are rebuilt from the caller's
 * kernels on every forward.  With the cache on, images are built once per distinct (weight pointers,
 * shape, tiling) and reused; the caller must call rgnn_weight_cache_clear() after changing cached weights
 * in place.  A cache miss inside CUDA-graph capture is an error (run the layer once eagerly first).
 * The key holds device ADDRESSES, not contents: a caller that frees a cached weight and later receives the same address for a
 * different weight (a new model in the same process) must call rgnn_weight_cache_clear() in between, or the old images are
 * used silently.  The Python binding does this itself (it tracks the owning tensor's lifetime and version per address). */
RGNN_API int rgnn_set_weight_cache(int enable);
RGNN_API int rgnn_weight_cache_clear(void);

/* Upper bound of the scratch a forward of `layer_kind` needs.  mlp_layers = number of kernels
 * in the widest edge/aggregation MLP (0 if none). */
RGNN_API size_t rgnn_workspace_bytes(const rgnn_plan_t* plan, int layer_kind, int32_t d_in, int32_t d_out,
                            int32_t mlp_layers);

/* ---- gnns/rgcn.py:8-117  sparse_rgcn_layer ------------------------------------------------
 * edge_weights: host array of L device pointers, kernel [d_in * (1 + use_both), d_out].
 * num_incoming: [L, V] fp32 or NULL when normalize == 0.  num_timesteps > 1 needs d_in == d_out. */
RGNN_API int rgnn_rgcn_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                      const float* const* edge_weights, const float* num_incoming,
                      int activation, int aggregation, int normalize_by_num_incoming,
                      int use_both_source_and_target, int num_timesteps,
                      float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of ONE timestep of sparse_rgcn_layer with source-only messages (use_both_source_and_target = 0):
 * the gradients TensorFlow autodiff produces for gnns/rgcn.py:84-114 (models/sparse_graph_model.py:253-260).
 *   out       forward output [V, d_out] (act' is evaluated from it; gelu recomputes the pre-activation)
 *   grad_out  dLoss/d out [V, d_out]
 *   grad_node_embeddings [V, d_in] or NULL; grad_edge_weights: host array of L device pointers [d_in, d_out] or NULL
 * 'max' aggregation is not differentiated in this build (RGNN_E_UNSUPPORTED).  The first backward on a plan builds
 * a reverse (by-source) index inside it (not thread-safe).  Workspace: rgnn_workspace_bytes(RGNN_LAYER_RGCN_BACKWARD). */
RGNN_API int rgnn_rgcn_backward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                       const float* const* edge_weights, const float* num_incoming,
                       int activation, int aggregation, int normalize_by_num_incoming,
                       const float* out, const float* grad_out,
                       float* grad_node_embeddings, float* const* grad_edge_weights,
                       void* workspace, size_t workspace_bytes, void* stream);

/* graph_num_layers x sparse_rgcn_layer in one call -- the GNN loop of
 * Sparse_Graph_Model.__build_graph_propagation_model (models/sparse_graph_model.py:176-191) without the scaffold's
 * dropout / residual / inter-layer Dense.  edge_weights: host array of num_layers * L pointers (layer-major),
 * every kernel [d, d]; the layers share activation / aggregation / normalisation like RGCN_Model does. */
RGNN_API int rgnn_rgcn_stack_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d, int32_t num_layers,
                            const float* const* edge_weights, const float* num_incoming,
                            int activation, int aggregation, int normalize_by_num_incoming,
                            float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gnns/ggnn.py:8-95  sparse_ggnn_layer --------------------------------------------------
 * cell_kernel [d, 3d] (GRU, gates z|r|h) or [d, d] (RNN); cell_recurrent_kernel same shape;
 * cell_bias [3d] / [d].  d_in must equal d_out (the cell state is the node state: ggnn.py:92). */
RGNN_API int rgnn_ggnn_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                      const float* const* edge_weights, const float* cell_kernel,
                      const float* cell_recurrent_kernel, const float* cell_bias,
                      int cell_kind, int activation, int aggregation, int num_timesteps,
                      float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gnns/rgat.py:9-141  sparse_rgat_layer -------------------------------------------------
 * attention: host array of L device pointers [2 * d_out]; head k uses [k*2d, (k+1)*2d),
 * first d for the source, next d for the target (rgat.py:110-111), d = d_out / num_heads. */
RGNN_API int rgnn_rgat_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                      const float* const* edge_weights, const float* const* attention,
                      int num_heads, int activation, int num_timesteps,
                      float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gnns/gnn_film.py:8-122  sparse_gnn_film_layer -----------------------------------------
 * film_weights: L pointers, kernel [d_in, 2*d_out] (gamma = cols [0,d), beta = cols [d,2d)).
 * ln_gamma / ln_beta: [num_timesteps, d_out] (one LayerNorm scope per timestep: gnn_film.py:120). */
RGNN_API int rgnn_film_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                      const float* const* edge_weights, const float* const* film_weights,
                      const float* num_incoming, const float* ln_gamma, const float* ln_beta,
                      int activation, int aggregation, int normalize_by_num_incoming, int num_timesteps,
                      float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gnns/gnn_edge_mlp.py:7-122  sparse_gnn_edge_mlp_layer ---------------------------------
 * mlp_kernels: host array of L * (num_edge_hidden_layers + 1) device pointers, type-major;
 * layer j of every type has shape [mlp_dims[j], mlp_dims[j+1]]; mlp_dims host [layers + 1],
 * mlp_dims[0] = d_in * (1 + use_target_state_as_input), mlp_dims[last] = d_out.
 * Hidden activation is ELU regardless of `activation` (gnn_edge_mlp.py:76). */
RGNN_API int rgnn_edge_mlp_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                          const float* const* mlp_kernels, const int32_t* mlp_dims, int num_edge_hidden_layers,
                          const float* num_incoming, const float* ln_gamma, const float* ln_beta,
                          int activation, int aggregation, int normalize_by_num_incoming,
                          int use_target_state_as_input, int num_timesteps,
                          float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gnns/rgin.py:7-142  sparse_rgin_layer -------------------------------------------------
 * num_edge_mlp_hidden_layers < 0  <=>  None (messages are the raw source states, no activation).
 * num_aggr_mlp_hidden_layers < 0  <=>  None.  aggr_kernels: host array of (layers+1) pointers,
 * aggr_dims host [layers + 2].  Edge-MLP hidden activation = `activation` (rgin.py:95). */
RGNN_API int rgnn_rgin_forward(const rgnn_plan_t* plan, const float* node_embeddings, int32_t d_in, int32_t d_out,
                      const float* const* edge_mlp_kernels, const int32_t* edge_mlp_dims,
                      int num_edge_mlp_hidden_layers,
                      const float* const* aggr_kernels, const int32_t* aggr_dims,
                      int num_aggr_mlp_hidden_layers,
                      const float* ln_gamma, const float* ln_beta,
                      int activation, int aggregation, int use_target_state_as_input, int num_timesteps,
                      float* out, void* workspace, size_t workspace_bytes, void* stream);

/* gnns/rgdcn.py:8-171 -- relational graph DYNAMIC convolution: the state is split into num_channels channels of
 * K = d / num_channels; message of edge (u -> v, type l), channel c:  h_u[c,:] . W[v,l,c]  with the K x K kernel
 * W[v,l,c] = reshape(act(F_{l,c} . x_v)), x_v = h_v (use_full_state) or h_v[c,:]; then 1/(c+1e-7) scaling, aggregation over
 * all types, activation.  channel_weights: host array of L*num_channels device pointers, entry l*num_channels + c =
 * F_{l,c} [d or K, K*K] (Keras Dense kernel; tie_channel_weights = the same pointer for every c).  K must be a power of
 * two in [4, 128], d <= 512.  For sum / mean / sqrt_n the per-(target, type) source rows are summed before ONE matvec per
 * channel (the kernel depends on the target only); max applies it per edge. */
RGNN_API int rgnn_rgdcn_forward(const rgnn_plan_t* plan, const float* h, int32_t d, int32_t num_channels,
                       const float* const* channel_weights, int use_full_state, const float* num_incoming,
                       int activation, int aggregation, int normalize, int num_timesteps, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- one large graph over several GPUs: node-range partition + halo exchange --------------------------
 * The reference is single-device; this is the multi-GPU form of ITS batch (tasks/varmisuse_task.py:451-538 packs up to
 * 100k nodes per batch): rank r of `world` owns the nodes [cuts[r], cuts[r+1]) and every edge whose TARGET it owns, so the
 * segment reductions / softmax / layer norm / GRU of every layer stay local; before each layer the states of the remote
 * SOURCE nodes ("halo") are refreshed.  One process per GPU; weights are replicated.
 *
 * rgnn_halo_plan_create: adjacency_lists hold GLOBAL node ids (device, int32 [E_l, 2]); edges whose target another rank
 * owns are dropped, so every rank may pass the same lists or only its own shard.  Built on the device (order-preserving
 * select, radix sort + unique of the remote sources, renumbering).  Local numbering: owned node g -> g - cuts[rank];
 * halo nodes follow, sorted by global id (hence grouped by owner).  rgnn_halo_plan_graph() is the ordinary plan over the
 * local ids with num_targets = the owned rows: pass it to any rgnn_<x>_forward with node states [n_own + n_halo, d]
 * (num_timesteps = 1 per call; call rgnn_halo_exchange between steps).  Synchronises the stream (it sizes buffers). */
RGNN_API int rgnn_halo_plan_create(rgnn_halo_plan_t** out, int32_t rank, int32_t world, const int64_t* cuts /* host [world+1] */,
                          int32_t num_edge_types, const int32_t* const* adjacency_lists, const int64_t* num_edges,
                          void* stream);
RGNN_API int rgnn_halo_plan_destroy(rgnn_halo_plan_t* plan);
RGNN_API int32_t rgnn_halo_plan_num_own(const rgnn_halo_plan_t* plan);
RGNN_API int32_t rgnn_halo_plan_num_halo(const rgnn_halo_plan_t* plan);
RGNN_API int64_t rgnn_halo_plan_num_edges(const rgnn_halo_plan_t* plan, int32_t edge_type);   /* kept edges of one type */
RGNN_API rgnn_plan_t* rgnn_halo_plan_graph(rgnn_halo_plan_t* plan);                            /* owned by the halo plan */
/* Copies into caller DEVICE buffers (any may be NULL): halo_global / halo_owner / halo_row [n_halo] (global id, owning rank,
 * row inside the owner's state buffer) and the renumbered adjacency lists (host array of L device pointers, [E_l, 2]). */
RGNN_API int rgnn_halo_plan_export(const rgnn_halo_plan_t* plan, int32_t* halo_global, int32_t* halo_owner, int32_t* halo_row,
                          int32_t* const* local_adjacency_lists, void* stream);
/* Peer memory.  Every rank keeps TWO state buffers (layer t reads buffer t % 2 and writes its owned rows into the other
 * one) of [n_own + n_halo, d] floats, owned rows first, plus one flag array uint32[world], in memory that the other ranks
 * have mapped into their address space (rgnn_peer_* below, or any other mechanism: these are plain device pointers).
 * peer_states0/1 and peer_flags are host arrays of `world` device pointers valid IN THIS PROCESS; entry [rank] is this
 * rank's own memory.  The flag arrays must start zeroed. */
RGNN_API int rgnn_halo_plan_attach(rgnn_halo_plan_t* plan, void* const* peer_states0, void* const* peer_states1,
                          void* const* peer_flags);
/* Refresh rows [n_own, n_own + n_halo) of this rank's state buffer `buffer` with the owners' current rows, read straight
 * out of the owners' buffers over NVLink (one kernel, pull-based: no packing, no send side).  Collective: every rank calls
 * it the same number of times, in the same order; the kernel contains the cross-rank barrier ("every rank's owned rows of
 * this buffer are final") as system-scope flags, and keeps its epoch on the device, so a captured CUDA graph can be
 * replayed.  A peer that never arrives faults the kernel after 10 s instead of hanging the GPU. */
RGNN_API int rgnn_halo_exchange(rgnn_halo_plan_t* plan, int buffer, int32_t d, void* stream);
/* The same exchange off the caller's critical path: the pull is forked onto a stream of the plan (ordered after everything
 * enqueued on `stream` so far) and is NOT joined here.  The next rgnn_<x>_forward on rgnn_halo_plan_graph() joins it right
 * before its first kernel that reads halo rows -- GNN-FiLM runs its target-side gamma / beta GEMM (owned rows only) first, so
 * that GEMM overlaps the transfer over NVLink.  Capturable into a CUDA graph (fork / join through events). */
RGNN_API int rgnn_halo_exchange_overlapped(rgnn_halo_plan_t* plan, int buffer, int32_t d, void* stream);

/* CUDA-IPC plumbing for the above (one node): allocate zeroed device memory that peers can map, export its 64-byte handle
 * (ship it with any host-side channel, e.g. torch.distributed.all_gather_object), map a peer's allocation. */
RGNN_API int rgnn_peer_alloc(void** ptr, size_t bytes, void* handle_out /* RGNN_PEER_HANDLE_BYTES */);
RGNN_API int rgnn_peer_open(const void* handle, void** ptr);
RGNN_API int rgnn_peer_close(void* ptr);
RGNN_API int rgnn_peer_free(void* ptr);

/* ---- building blocks exported for tests / other hosts ---------------------------------------
 * utils/utils.py:23-33: aggregate `data` [M, d] (rows in the ORIGINAL type-major message order)
 * to [V, d] with the plan's segments -- the tf.unsorted_segment_<agg> call of rgcn.py:110. */
RGNN_API int rgnn_segment_aggregate(const rgnn_plan_t* plan, const float* data, int32_t d, int aggregation,
                           float* out, void* stream);
/* The edge stage on per-node transformed states: out[v,:] = agg over the incoming edges (u,v) of every type l of
 * s * table[u,l,:], table [V, L, d] row-major, s = 1/(num_incoming[l,v] + 1e-7) or 1 when num_incoming is NULL
 * (gnns/rgcn.py:84-112 and ggnn.py:76-90 with the per-type Dense applied per node first).  The backward gives
 * d_table[u,l,:] = sum over the outgoing edges (u,v) of type l of s * grad_out[v,:] / div(v) for sum / mean / sqrt_n
 * (RGNN_E_UNSUPPORTED for max); the reverse index is built inside the plan on first use. */
RGNN_API int rgnn_edge_aggregate_forward(const rgnn_plan_t* plan, const float* table, int32_t d, const float* num_incoming,
                                int aggregation, float* out, void* stream);
RGNN_API int rgnn_edge_aggregate_backward(const rgnn_plan_t* plan, const float* grad_out, int32_t d,
                                 const float* num_incoming, int aggregation, float* d_table, void* stream);
/* C[M,N] = act(A[M,K] . B[K,N] + bias) on the tensor cores with 3xTF32 split accumulation
 * (fp32-accurate); the node-level Dense of every layer (A.1). bias may be NULL.  workspace: caller scratch of at least
 * rgnn_dense_workspace_bytes(m, k, n) bytes (weight images; split-K partial tiles of the backward) -- the library
 * allocates nothing per call. */
RGNN_API size_t rgnn_dense_workspace_bytes(int32_t m, int32_t k, int32_t n);
RGNN_API int rgnn_dense_forward(const float* a, int32_t m, int32_t k, const float* b, int32_t n,
                       const float* bias, int activation, float* c, void* workspace, size_t workspace_bytes, void* stream);
/* Gradients of the linear map C = A . B: grad_a [M,K] = grad_c . B^T, grad_b [K,N] = A^T . grad_c (split-K on the tensor
 * cores, deterministic).  Either output may be NULL.  The reference gets these from TF autodiff of its Dense kernels
 * (models/sparse_graph_model.py:253-260). */
RGNN_API int rgnn_dense_backward(const float* a, int32_t m, int32_t k, const float* b, int32_t n, const float* grad_c,
                        float* grad_a, float* grad_b, void* workspace, size_t workspace_bytes, void* stream);
/* tf.contrib.layers.layer_norm over the last axis, eps 1e-12 (A.5). */
RGNN_API int rgnn_layer_norm(const float* x, int32_t rows, int32_t d, const float* gamma, const float* beta,
                    float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGNN_H_ */
