"""numpy restatement of the reference GNN layers (TEST INFRASTRUCTURE -- see oracle/__init__.py).

PINNED AGAINST THE REFERENCE'S OWN CODE: the reference has no tests or golden vectors and TF1 is not installable here, but
its layer functions are plain Python over ~25 tf.* calls, so tests/golden/make_ref_fixtures.py EXECUTES the unmodified
/root/reference/gnns/*.py + utils/utils.py through a numpy-backed ``tensorflow`` / ``dpu_utils`` stand-in (tests/tf1_shim)
and commits the outputs as tests/golden/ref_*.npz; tests/test_reference_pin.py holds this oracle to them at 1e-12
(17 small cases covering every keyword argument that changes the op order, and BASELINE.json configs 2-5 at full size).
What remains an assumption is only the TF 1.13 / Keras / dpu_utils KERNEL semantics the stand-in restates (SURVEY.md
Appendix A); tests/golden/make_tf1_fixtures.py discharges it on a machine with the real stack.

Every function follows the reference op order literally (gather -> per-edge
transform -> scale -> concat -> unsorted segment reduce -> activation), so that
run with ``dtype=np.float32`` it is the stand-in for "the reference TF1 CPU
path" and with ``dtype=np.float64`` it is the accuracy truth.

Citations are into /root/reference (microsoft/tf-gnn-samples @ ff14b96a).

Weight containers are plain dicts of numpy arrays, Keras orientation
(``kernel[in, out]``, ``y = x @ kernel``):
  rgcn      {"edge_weights": L x [Din*(1|2), D]}
  ggnn      {"edge_weights": L x [Din, D], "cell": {"kernel", "recurrent_kernel", "bias"}}
  rgat      {"edge_weights": L x [Din, D], "attention": L x [2D]}
  film      {"edge_weights": L x [Din, D], "film_weights": L x [Din, 2D], "ln_gamma", "ln_beta"}
  edge_mlp  {"edge_mlps": L x [kernels...], "ln_gamma", "ln_beta"}
  rgin      {"edge_mlps": L x [kernels...] | None, "aggr_mlp": [kernels...] | None, "ln_gamma", "ln_beta"}
``ln_gamma`` / ``ln_beta`` are either one [D] vector (shared) or a list with one
[D] vector per timestep (the reference creates a fresh LayerNorm scope per
timestep: gnn_film.py:120, gnn_edge_mlp.py:119, rgin.py:139).
"""
from typing import Callable, Dict, List, Optional, Sequence

import math
import numpy as np

BIG_NUMBER = 1e7      # utils/utils.py:6
SMALL_NUMBER = 1e-7   # utils/utils.py:7

_F32_LOWEST = float(np.finfo(np.float32).min)


# --------------------------------------------------------------------------------------
# utils/utils.py:36-58  get_activation
# --------------------------------------------------------------------------------------
def _erf(x: np.ndarray) -> np.ndarray:
    out = np.empty_like(x)
    flat_in, flat_out = x.reshape(-1), out.reshape(-1)
    for i in range(flat_in.shape[0]):           # only used on small tensors / via vectorize below
        flat_out[i] = math.erf(float(flat_in[i]))
    return out


try:                                            # scipy is in the image; keep a pure fallback
    from scipy.special import erf as _sp_erf

    def _erf(x):                                # noqa: F811
        return _sp_erf(x)
except Exception:                               # pragma: no cover
    pass


def get_activation(activation_fun: Optional[str]) -> Optional[Callable[[np.ndarray], np.ndarray]]:
    """utils/utils.py:36-58.  Returns None for None/'linear' exactly like the reference."""
    if activation_fun is None:
        return None
    name = activation_fun.lower()
    if name == 'linear':
        return None
    if name == 'tanh':
        return np.tanh
    if name == 'relu':
        return lambda x: np.maximum(x, x.dtype.type(0))
    if name == 'leaky_relu':                    # tf.nn.leaky_relu default alpha=0.2
        return lambda x: np.where(x > 0, x, x * x.dtype.type(0.2))
    if name == 'elu':
        return lambda x: np.where(x > 0, x, np.expm1(np.minimum(x, x.dtype.type(0))))
    if name == 'selu':
        scale, alpha = 1.0507009873554805, 1.6732632423543772
        return lambda x: x.dtype.type(scale) * np.where(
            x > 0, x, x.dtype.type(alpha) * np.expm1(np.minimum(x, x.dtype.type(0))))
    if name == 'gelu':                          # utils/utils.py:52-56 (exact erf form)
        def gelu(x):
            cdf = x.dtype.type(0.5) * (x.dtype.type(1.0) + _erf(x / x.dtype.type(math.sqrt(2.0)))).astype(x.dtype)
            return x * cdf
        return gelu
    raise ValueError("Unknown activation function '%s'!" % activation_fun)


def _apply_act(fn, x):
    # The reference would raise TypeError calling None; the product treats None as identity
    # (documented superset).  The oracle does the same so both can be compared.
    return x if fn is None else fn(x)


# --------------------------------------------------------------------------------------
# utils/utils.py:23-33  get_aggregation_function  (tf.unsorted_segment_*)
# --------------------------------------------------------------------------------------
def unsorted_segment_sum(data: np.ndarray, segment_ids: np.ndarray, num_segments: int) -> np.ndarray:
    out = np.zeros((num_segments,) + data.shape[1:], dtype=data.dtype)
    np.add.at(out, segment_ids, data)           # sequential accumulation in message order (TF CPU kernel)
    return out


def _segment_count(segment_ids: np.ndarray, num_segments: int, dtype) -> np.ndarray:
    n = np.bincount(segment_ids, minlength=num_segments).astype(dtype)
    return np.maximum(n, dtype(1))              # TF clamps the divisor at 1 for empty segments


def unsorted_segment_mean(data, segment_ids, num_segments):
    s = unsorted_segment_sum(data, segment_ids, num_segments)
    n = _segment_count(segment_ids, num_segments, data.dtype.type)
    return s / n.reshape((-1,) + (1,) * (data.ndim - 1))


def unsorted_segment_sqrt_n(data, segment_ids, num_segments):
    s = unsorted_segment_sum(data, segment_ids, num_segments)
    n = _segment_count(segment_ids, num_segments, data.dtype.type)
    return s / np.sqrt(n).reshape((-1,) + (1,) * (data.ndim - 1))


def unsorted_segment_max(data, segment_ids, num_segments):
    # empty segment -> numeric_limits<float>::lowest()  (float32 lowest: the reference runs in fp32)
    out = np.full((num_segments,) + data.shape[1:], _F32_LOWEST, dtype=data.dtype)
    np.maximum.at(out, segment_ids, data)
    return out


def get_aggregation_function(aggregation_fun: Optional[str]):
    """utils/utils.py:23-33."""
    if aggregation_fun in ['sum', 'unsorted_segment_sum']:
        return unsorted_segment_sum
    if aggregation_fun in ['max', 'unsorted_segment_max']:
        return unsorted_segment_max
    if aggregation_fun in ['mean', 'unsorted_segment_mean']:
        return unsorted_segment_mean
    if aggregation_fun in ['sqrt_n', 'unsorted_segment_sqrt_n']:
        return unsorted_segment_sqrt_n
    raise ValueError("Unknown aggregation function '%s'!" % aggregation_fun)


# --------------------------------------------------------------------------------------
# third-party defaults
# --------------------------------------------------------------------------------------
def dense(x: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """tf.keras.layers.Dense(use_bias=False): y = x @ kernel, kernel [in, out] (A.1)."""
    return x @ kernel.astype(x.dtype)


def hard_sigmoid(x: np.ndarray) -> np.ndarray:
    """Keras hard_sigmoid: clip(0.2*x + 0.5, 0, 1) -- TF1 GRUCell recurrent_activation default (A.4)."""
    t = x.dtype.type
    return np.clip(t(0.2) * x + t(0.5), t(0), t(1))


def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """tf.contrib.layers.layer_norm defaults (A.5): last axis, biased variance, eps 1e-12,
    evaluated like tf.nn.batch_normalization: x*inv + (beta - mean*inv), inv = rsqrt(var+eps)*gamma."""
    t = x.dtype.type
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    inv = (t(1) / np.sqrt(var + t(eps))) * gamma.astype(x.dtype)
    return x * inv + (beta.astype(x.dtype) - mean * inv)


def _ln_params(weights: Dict, t: int, dim: int, dtype):
    g, b = weights.get("ln_gamma"), weights.get("ln_beta")
    if g is None:
        g = np.ones(dim, dtype)
    if b is None:
        b = np.zeros(dim, dtype)
    if isinstance(g, (list, tuple)):
        g = g[t]
    if isinstance(b, (list, tuple)):
        b = b[t]
    return np.asarray(g), np.asarray(b)


def mlp_apply(kernels: Sequence[np.ndarray], x: np.ndarray, hidden_act) -> np.ndarray:
    """utils/utils.py:77-126  MLP: bias-free Dense stack, activation on hidden layers only,
    linear output layer; dropout rate 0.0 == identity."""
    a = x
    for k in kernels[:-1]:
        a = _apply_act(hidden_act, dense(a, k))
    return dense(a, kernels[-1])


def gru_cell(x, h, kernel, recurrent_kernel, bias, act):
    """tf.keras.layers.GRUCell (TF 1.13 defaults: hard_sigmoid, reset_after=False), gate order z,r,h (A.4)."""
    d = h.shape[1]
    kernel, recurrent_kernel, bias = (a.astype(x.dtype) for a in (kernel, recurrent_kernel, bias))
    xz = x @ kernel[:, :d] + bias[:d]
    xr = x @ kernel[:, d:2 * d] + bias[d:2 * d]
    xh = x @ kernel[:, 2 * d:] + bias[2 * d:]
    z = hard_sigmoid(xz + h @ recurrent_kernel[:, :d])
    r = hard_sigmoid(xr + h @ recurrent_kernel[:, d:2 * d])
    hh = _apply_act(act, xh + (r * h) @ recurrent_kernel[:, 2 * d:])
    return z * h + (x.dtype.type(1) - z) * hh


def rnn_cell(x, h, kernel, recurrent_kernel, bias, act):
    """tf.keras.layers.SimpleRNNCell: act(x@W + b + h@U) (A.4)."""
    kernel, recurrent_kernel, bias = (a.astype(x.dtype) for a in (kernel, recurrent_kernel, bias))
    return _apply_act(act, (x @ kernel + bias) + h @ recurrent_kernel)


def get_gated_unit(gated_unit: str, activation_function: Optional[str]):
    """utils/utils.py:10-20 (LSTM is unusable as the reference calls it -- one state: not implemented)."""
    act = get_activation(activation_function)
    name = gated_unit.lower()
    if name == 'rnn':
        return lambda x, h, c: rnn_cell(x, h, c["kernel"], c["recurrent_kernel"], c["bias"], act)
    if name == 'gru':
        return lambda x, h, c: gru_cell(x, h, c["kernel"], c["recurrent_kernel"], c["bias"], act)
    if name == 'lstm':
        raise NotImplementedError("LSTMCell needs [h, c]; the reference passes one state (ggnn.py:92) and fails")
    raise Exception("Unknown RNN cell type '%s'." % gated_unit)


def unsorted_segment_log_softmax(logits, segment_ids, num_segments):
    """dpu_utils.tfutils.unsorted_segment_log_softmax (A.7)."""
    m = unsorted_segment_max(logits, segment_ids, num_segments)
    z = logits - m[segment_ids]
    s = unsorted_segment_sum(np.exp(z), segment_ids, num_segments)
    with np.errstate(divide="ignore"):          # empty segments have s = 0; their log is never gathered
        log_s = np.log(s)
    return z - log_s[segment_ids]


def _prep(node_embeddings, adjacency_lists, dtype):
    h = np.asarray(node_embeddings).astype(dtype)
    adj = [np.asarray(a).reshape(-1, 2).astype(np.int64) for a in adjacency_lists]
    return h, adj


# --------------------------------------------------------------------------------------
# gnns/rgcn.py:8-117
# --------------------------------------------------------------------------------------
def sparse_rgcn_layer(node_embeddings, adjacency_lists, type_to_num_incoming_edges, state_dim,
                      num_timesteps=1, activation_function="tanh", message_aggregation_function="sum",
                      normalize_by_num_incoming=True, use_both_source_and_target=False,
                      *, weights: Dict, dtype=np.float64):
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]                                                    # rgcn.py:67
    activation_fn = get_activation(activation_function)                        # :72
    aggregation_fn = get_aggregation_function(message_aggregation_function)   # :73
    message_targets = np.concatenate([a[:, 1] for a in adj]) if adj else np.zeros(0, np.int64)   # :78
    cnt = np.asarray(type_to_num_incoming_edges).astype(dtype) if type_to_num_incoming_edges is not None else None
    cur = h
    for _ in range(num_timesteps):                                            # :81
        messages_per_type = []
        for l, a in enumerate(adj):                                           # :84
            src, tgt = a[:, 0], a[:, 1]                                       # :85-86
            s = cur[src]                                                      # :88
            if use_both_source_and_target:                                    # :91-96
                s = np.concatenate([s, cur[tgt]], axis=-1)
            msg = dense(s, weights["edge_weights"][l])                        # :98
            if normalize_by_num_incoming:                                     # :100-104
                c = cnt[l][tgt]
                msg = (dtype(1.0) / (c + dtype(SMALL_NUMBER)))[:, None] * msg
            messages_per_type.append(msg)
        out_dim = weights["edge_weights"][0].shape[1] if len(adj) else state_dim
        cur_messages = np.concatenate(messages_per_type, axis=0) if adj else np.zeros((0, out_dim), dtype)  # :108
        agg = aggregation_fn(cur_messages, message_targets, num_nodes)        # :110
        cur = _apply_act(activation_fn, agg)                                  # :114
    return cur


# --------------------------------------------------------------------------------------
# gnns/ggnn.py:8-95
# --------------------------------------------------------------------------------------
def sparse_ggnn_layer(node_embeddings, adjacency_lists, state_dim, num_timesteps=1,
                      gated_unit_type="gru", activation_function="tanh", message_aggregation_function="sum",
                      *, weights: Dict, dtype=np.float64):
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]
    aggregation_fn = get_aggregation_function(message_aggregation_function)   # ggnn.py:55
    cell = get_gated_unit(gated_unit_type, activation_function)               # :56
    message_targets = np.concatenate([a[:, 1] for a in adj])                  # :68
    cur = h
    for _ in range(num_timesteps):                                            # :71
        msgs = []
        for l, a in enumerate(adj):                                           # :76
            msgs.append(dense(cur[a[:, 0]], weights["edge_weights"][l]))      # :78-82
        msgs = np.concatenate(msgs, axis=0)                                   # :86
        agg = aggregation_fn(msgs, message_targets, num_nodes)                # :87-90
        cur = cell(agg, cur, weights["cell"])                                 # :92  (inputs=agg, state=h)
    return cur


# --------------------------------------------------------------------------------------
# gnns/rgat.py:9-141
# --------------------------------------------------------------------------------------
def sparse_rgat_layer(node_embeddings, adjacency_lists, state_dim, num_heads=4, num_timesteps=1,
                      activation_function="tanh", *, weights: Dict, dtype=np.float64):
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]
    D = weights["edge_weights"][0].shape[1]
    per_head_dim = D // num_heads                                             # rgat.py:61
    activation_fn = get_activation(activation_function)
    message_targets = np.concatenate([a[:, 1] for a in adj])                  # :80
    leaky = get_activation("leaky_relu")
    cur = h
    for _ in range(num_timesteps):                                            # :83
        per_head_msgs, per_head_coeffs = [], []
        for l, a in enumerate(adj):                                           # :91
            src, tgt = a[:, 0], a[:, 1]
            transformed = dense(cur, weights["edge_weights"][l])              # :95-96 (nodes first)
            ts = transformed[src].reshape(-1, num_heads, per_head_dim)        # :98-104
            tt = transformed[tgt].reshape(-1, num_heads, per_head_dim)
            both = np.concatenate([ts, tt], axis=-1)                          # :106-109  [E,K,2d]
            pars = np.asarray(weights["attention"][l]).astype(dtype).reshape(num_heads, 2 * per_head_dim)  # :110-111
            coeff = leaky(np.einsum('vki,ki->vk', both, pars))                # :112-115
            per_head_msgs.append(ts)
            per_head_coeffs.append(coeff)
        per_head_msgs = np.concatenate(per_head_msgs, axis=0)                 # :120
        per_head_coeffs = np.concatenate(per_head_coeffs, axis=0)             # :121
        heads = []
        for k in range(num_heads):                                            # :124
            att = np.exp(unsorted_segment_log_softmax(per_head_coeffs[:, k], message_targets, num_nodes))  # :126-130
            heads.append(unsorted_segment_sum(att[:, None] * per_head_msgs[:, k, :], message_targets, num_nodes))  # :131-136
        cur = _apply_act(activation_fn, np.concatenate(heads, axis=-1))       # :138
    return cur


# --------------------------------------------------------------------------------------
# gnns/gnn_film.py:8-122
# --------------------------------------------------------------------------------------
def sparse_gnn_film_layer(node_embeddings, adjacency_lists, type_to_num_incoming_edges, state_dim,
                          num_timesteps=1, activation_function="ReLU", message_aggregation_function="sum",
                          normalize_by_num_incoming=False, *, weights: Dict, dtype=np.float64):
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]
    D = weights["edge_weights"][0].shape[1]
    activation_fn = get_activation(activation_function)
    aggregation_fn = get_aggregation_function(message_aggregation_function)
    message_targets = np.concatenate([a[:, 1] for a in adj])                  # gnn_film.py:82
    cnt = np.asarray(type_to_num_incoming_edges).astype(dtype) if type_to_num_incoming_edges is not None else None
    cur = h
    for t in range(num_timesteps):                                            # :85
        per_type = []
        for l, a in enumerate(adj):                                           # :88
            src, tgt = a[:, 0], a[:, 1]
            msg = dense(cur[src], weights["edge_weights"][l])                 # :91-94
            if normalize_by_num_incoming:                                     # :96-100
                msg = (dtype(1.0) / (cnt[l][tgt] + dtype(SMALL_NUMBER)))[:, None] * msg
            film = dense(cur, weights["film_weights"][l])                     # :102
            pm = film[tgt]                                                    # :103-104
            gamma, beta = pm[:, :D], pm[:, D:]                                # :105-106
            per_type.append(gamma * msg + beta)                               # :108
        allm = _apply_act(activation_fn, np.concatenate(per_type, axis=0))    # :111-112 (act before the sum)
        agg = aggregation_fn(allm, message_targets, num_nodes)                # :113-116
        g, b = _ln_params(weights, t, D, dtype)
        cur = layer_norm(agg, g, b)                                           # :120 (no activation after: :118)
    return cur


# --------------------------------------------------------------------------------------
# gnns/gnn_edge_mlp.py:7-122
# --------------------------------------------------------------------------------------
def sparse_gnn_edge_mlp_layer(node_embeddings, adjacency_lists, type_to_num_incoming_edges, state_dim,
                              num_timesteps=1, activation_function="ReLU", message_aggregation_function="sum",
                              normalize_by_num_incoming=False, use_target_state_as_input=True,
                              num_edge_hidden_layers=1, *, weights: Dict, dtype=np.float64):
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]
    activation_fn = get_activation(activation_function)
    aggregation_fn = get_aggregation_function(message_aggregation_function)
    elu = get_activation("elu")                                               # gnn_edge_mlp.py:76 (hard-coded)
    message_targets = np.concatenate([a[:, 1] for a in adj])                  # :81
    cnt = np.asarray(type_to_num_incoming_edges).astype(dtype) if type_to_num_incoming_edges is not None else None
    D = weights["edge_mlps"][0][-1].shape[1]
    cur = h
    for t in range(num_timesteps):                                            # :84
        per_type = []
        for l, a in enumerate(adj):                                           # :87
            src, tgt = a[:, 0], a[:, 1]
            x = cur[src]                                                      # :90-92
            if use_target_state_as_input:                                     # :95-100
                x = np.concatenate([x, cur[tgt]], axis=1)
            assert len(weights["edge_mlps"][l]) == num_edge_hidden_layers + 1
            msg = mlp_apply(weights["edge_mlps"][l], x, elu)                  # :102
            if normalize_by_num_incoming:                                     # :104-108
                msg = (dtype(1.0) / (cnt[l][tgt] + dtype(SMALL_NUMBER)))[:, None] * msg
            per_type.append(msg)
        allm = _apply_act(activation_fn, np.concatenate(per_type, axis=0))    # :111-112
        agg = aggregation_fn(allm, message_targets, num_nodes)                # :113-116
        g, b = _ln_params(weights, t, D, dtype)
        cur = layer_norm(agg, g, b)                                           # :119
    return cur


# --------------------------------------------------------------------------------------
# gnns/rgin.py:7-142
# --------------------------------------------------------------------------------------
def sparse_rgin_layer(node_embeddings, adjacency_lists, state_dim, num_timesteps=1,
                      activation_function="ReLU", message_aggregation_function="sum",
                      use_target_state_as_input=False, num_edge_MLP_hidden_layers=1,
                      num_aggr_MLP_hidden_layers=None, *, weights: Dict, dtype=np.float64):
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]
    activation_fn = get_activation(activation_function)
    aggregation_fn = get_aggregation_function(message_aggregation_function)
    edge_mlps = weights.get("edge_mlps") if num_edge_MLP_hidden_layers is not None else None   # rgin.py:86-89
    aggr_mlp = weights.get("aggr_mlp") if num_aggr_MLP_hidden_layers is not None else None     # :78-84
    message_targets = np.concatenate([a[:, 1] for a in adj])                  # :100
    cur = h
    for t in range(num_timesteps):                                            # :103
        per_type = []
        for l, a in enumerate(adj):                                           # :106
            src, tgt = a[:, 0], a[:, 1]
            x = cur[src]                                                      # :109-111
            if use_target_state_as_input:                                     # :114-119
                x = np.concatenate([x, cur[tgt]], axis=1)
            if edge_mlps is not None:                                         # :121-124
                x = mlp_apply(edge_mlps[l], x, activation_fn)
            per_type.append(x)
        allm = np.concatenate(per_type, axis=0)                               # :127
        if edge_mlps is not None:                                             # :128-129
            allm = _apply_act(activation_fn, allm)
        agg = aggregation_fn(allm, message_targets, num_nodes)                # :130-133
        new = agg
        if aggr_mlp is not None:                                              # :136-137
            new = mlp_apply(aggr_mlp, new, activation_fn)
        new = _apply_act(activation_fn, new)                                  # :138
        g, b = _ln_params(weights, t, new.shape[1], dtype)
        cur = layer_norm(new, g, b)                                           # :139
    return cur


# --------------------------------------------------------------------------------------
# gnns/rgdcn.py:8-171
# --------------------------------------------------------------------------------------
def sparse_rgdcn_layer(node_embeddings, adjacency_lists, type_to_num_incoming_edges, num_channels=8, channel_dim=16,
                       num_timesteps=1, use_full_state_for_channel_weights=False, tie_channel_weights=False,
                       activation_function="tanh", message_aggregation_function="sum", normalize_by_num_incoming=True,
                       *, weights: Dict, dtype=np.float64):
    """weights: {"channel_weights": L x C' x [D or K, K*K]}, C' = 1 when tied (rgdcn.py:96-107) else num_channels."""
    h, adj = _prep(node_embeddings, adjacency_lists, dtype)
    num_nodes = h.shape[0]                                                    # rgdcn.py:88
    activation_fn = get_activation(activation_function)                        # :92
    aggregation_fn = get_aggregation_function(message_aggregation_function)   # :93
    message_targets = np.concatenate([a[:, 1] for a in adj])                  # :113
    cnt = np.asarray(type_to_num_incoming_edges).astype(dtype) if type_to_num_incoming_edges is not None else None
    C, K = num_channels, channel_dim
    cur = h
    for _ in range(num_timesteps):                                            # :116
        chunked = cur.reshape(-1, C, K)                                       # :117-118
        new_chunks = []
        for c in range(C):                                                    # :121
            chan = chunked[:, c, :]                                           # :122
            per_type = []
            for l, a in enumerate(adj):                                       # :126
                src, tgt = a[:, 0], a[:, 1]
                src_states = chan[src]                                        # :129-131
                inp = cur if use_full_state_for_channel_weights else chan     # :133-136
                kernel = weights["channel_weights"][l][0 if tie_channel_weights else c]      # :139
                ew = _apply_act(activation_fn, dense(inp, kernel)).reshape(-1, K, K)       # :140-141 (Dense carries the activation)
                msgs = np.einsum('vi,vij->vj', src_states, ew[tgt])                        # :142-146
                if normalize_by_num_incoming:                                 # :147-151
                    msgs = (dtype(1.0) / (cnt[l][tgt] + dtype(SMALL_NUMBER)))[:, None] * msgs
                per_type.append(msgs)
            agg = aggregation_fn(np.concatenate(per_type, axis=0), message_targets, num_nodes)   # :155-159
            new_chunks.append(_apply_act(activation_fn, agg))                 # :160
        cur = np.concatenate(new_chunks, axis=1)                              # :164-165
    return cur


LAYERS = {
    "rgcn": sparse_rgcn_layer,
    "ggnn": sparse_ggnn_layer,
    "rgat": sparse_rgat_layer,
    "gnn-film": sparse_gnn_film_layer,
    "gnn-edge-mlp": sparse_gnn_edge_mlp_layer,
    "rgin": sparse_rgin_layer,
    "rgdcn": sparse_rgdcn_layer,
}


def max_norm_rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """Parity metric (SURVEY 7 'Tolerance definition'): maxabs(a-b)/maxabs(b)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = float(np.max(np.abs(b))) if b.size else 0.0
    if denom == 0.0:
        return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / denom)
