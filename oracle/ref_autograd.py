"""Differentiable restatement of the reference layers in torch float64 on the CPU (TEST INFRASTRUCTURE -- see
oracle/__init__.py; never imported by the product).

Pinning: the FORWARD of every function here equals oracle/ref_layers.py (tests/test_oracle_autograd.py), which is held at 1e-12
to the reference's own executed code (tests/test_reference_pin.py, tests/test_reference_fuzz.py); the GRADIENTS are then the
derivatives of a pinned function, checked by torch.autograd.gradcheck.  (TF's autodiff itself cannot run here -- there is no
reference-produced gradient to compare with.)

Purpose: the reference gets its gradients from TF autodiff of exactly these op graphs
(models/sparse_graph_model.py:253-260), so torch autograd over the same op order in float64 is the gradient truth
the engine's backward passes are compared with.  Op order follows the reference literally: per edge type
gather -> per-edge transform -> scale -> concat -> unsorted segment reduce -> activation (file:line cited inline,
into /root/reference).

Weight containers: the dictionaries of oracle/ref_layers.py with torch float64 tensors (requires_grad as needed).
"""
from typing import Dict, List, Optional, Sequence

import math
import torch

SMALL_NUMBER = 1e-7                       # utils/utils.py:7
_F32_LOWEST = -3.4028234663852886e38      # tf.unsorted_segment_max of an empty segment (fp32 lowest)


def get_activation(name: Optional[str]):
    """utils/utils.py:36-58."""
    if name is None or name.lower() == "linear":
        return lambda x: x
    n = name.lower()
    if n == "tanh":
        return torch.tanh
    if n == "relu":
        return torch.relu
    if n == "leaky_relu":
        return lambda x: torch.where(x > 0, x, 0.2 * x)
    if n == "elu":
        return lambda x: torch.where(x > 0, x, torch.expm1(torch.clamp(x, max=0.0)))
    if n == "selu":
        return lambda x: 1.0507009873554805 * torch.where(x > 0, x, 1.6732632423543772 * torch.expm1(torch.clamp(x, max=0.0)))
    if n == "gelu":
        return lambda x: x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    raise ValueError("Unknown activation function '%s'!" % name)


def segment_reduce(data: torch.Tensor, ids: torch.Tensor, num_segments: int, how: str) -> torch.Tensor:
    """utils/utils.py:23-33 -> tf.unsorted_segment_{sum,max,mean,sqrt_n}."""
    shape = (num_segments,) + tuple(data.shape[1:])
    if how in ("sum", "unsorted_segment_sum", "mean", "unsorted_segment_mean", "sqrt_n", "unsorted_segment_sqrt_n"):
        s = torch.zeros(shape, dtype=data.dtype).index_add(0, ids, data)
        if how in ("sum", "unsorted_segment_sum"):
            return s
        n = torch.clamp(torch.bincount(ids, minlength=num_segments).to(data.dtype), min=1.0)
        n = n.reshape((-1,) + (1,) * (data.dim() - 1))
        return s / n if how in ("mean", "unsorted_segment_mean") else s / torch.sqrt(n)
    if how in ("max", "unsorted_segment_max"):
        out = torch.full(shape, _F32_LOWEST, dtype=data.dtype)
        idx = ids.reshape((-1,) + (1,) * (data.dim() - 1)).expand_as(data)
        return out.scatter_reduce(0, idx, data, reduce="amax", include_self=True)
    raise ValueError("Unknown aggregation function '%s'!" % how)


def hard_sigmoid(x):
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def layer_norm(x, gamma, beta, eps: float = 1e-12):
    """tf.contrib.layers.layer_norm defaults: last axis, biased variance, eps 1e-12."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    inv = gamma / torch.sqrt(var + eps)
    return x * inv + (beta - mean * inv)


def _ln(weights: Dict, t: int, dim: int):
    g, b = weights.get("ln_gamma"), weights.get("ln_beta")
    g = torch.ones(dim, dtype=torch.float64) if g is None else (g[t] if isinstance(g, (list, tuple)) else g)
    b = torch.zeros(dim, dtype=torch.float64) if b is None else (b[t] if isinstance(b, (list, tuple)) else b)
    return g, b


def mlp_apply(kernels: Sequence[torch.Tensor], x, hidden_act):
    """utils/utils.py:77-126."""
    for k in kernels[:-1]:
        x = hidden_act(x @ k)
    return x @ kernels[-1]


def _adj(adjacency_lists) -> List[torch.Tensor]:
    return [torch.as_tensor(a).reshape(-1, 2).long() for a in adjacency_lists]


def _scale(cnt, l, tgt):
    return (1.0 / (cnt[l][tgt] + SMALL_NUMBER)).unsqueeze(-1)


def sparse_rgcn_layer(h, adjacency_lists, type_to_num_incoming_edges, num_timesteps=1, activation_function="tanh",
                      message_aggregation_function="sum", normalize_by_num_incoming=True,
                      use_both_source_and_target=False, *, weights: Dict):
    """gnns/rgcn.py:67-117."""
    adj, act, V = _adj(adjacency_lists), get_activation(activation_function), h.shape[0]
    targets = torch.cat([a[:, 1] for a in adj])
    cur = h
    for _ in range(num_timesteps):
        msgs = []
        for l, a in enumerate(adj):
            s = cur[a[:, 0]]
            if use_both_source_and_target:
                s = torch.cat([s, cur[a[:, 1]]], dim=-1)
            m = s @ weights["edge_weights"][l]
            if normalize_by_num_incoming:
                m = _scale(type_to_num_incoming_edges, l, a[:, 1]) * m
            msgs.append(m)
        cur = act(segment_reduce(torch.cat(msgs), targets, V, message_aggregation_function))
    return cur


def sparse_ggnn_layer(h, adjacency_lists, num_timesteps=1, gated_unit_type="gru", activation_function="tanh",
                      message_aggregation_function="sum", *, weights: Dict):
    """gnns/ggnn.py:50-95; Keras GRUCell / SimpleRNNCell TF 1.13 defaults (hard_sigmoid, reset_after=False, z|r|h)."""
    adj, act, V = _adj(adjacency_lists), get_activation(activation_function), h.shape[0]
    targets = torch.cat([a[:, 1] for a in adj])
    c = weights["cell"]
    K, R, B = c["kernel"], c["recurrent_kernel"], c["bias"]
    d = h.shape[1]
    kind = gated_unit_type.lower()
    if kind not in ("gru", "rnn"):
        raise Exception("Unknown RNN cell type '%s'." % gated_unit_type)
    cur = h
    for _ in range(num_timesteps):
        m = segment_reduce(torch.cat([cur[a[:, 0]] @ weights["edge_weights"][l] for l, a in enumerate(adj)]), targets, V,
                           message_aggregation_function)
        if kind == "rnn":
            cur = act((m @ K + B) + cur @ R)
        else:
            z = hard_sigmoid(m @ K[:, :d] + B[:d] + cur @ R[:, :d])
            r = hard_sigmoid(m @ K[:, d:2 * d] + B[d:2 * d] + cur @ R[:, d:2 * d])
            hh = act(m @ K[:, 2 * d:] + B[2 * d:] + (r * cur) @ R[:, 2 * d:])
            cur = z * cur + (1.0 - z) * hh
    return cur


def sparse_rgat_layer(h, adjacency_lists, num_timesteps=1, num_heads=4, activation_function="tanh", *, weights: Dict):
    """gnns/rgat.py:58-141; dpu_utils unsorted_segment_log_softmax = x - max - log(sum(exp(x - max)))."""
    adj, act, V = _adj(adjacency_lists), get_activation(activation_function), h.shape[0]
    targets = torch.cat([a[:, 1] for a in adj])
    D = weights["edge_weights"][0].shape[1]
    dh = D // num_heads
    cur = h
    for _ in range(num_timesteps):
        msgs, coeffs = [], []
        for l, a in enumerate(adj):
            t = cur @ weights["edge_weights"][l]
            ts = t[a[:, 0]].reshape(-1, num_heads, dh)
            tt = t[a[:, 1]].reshape(-1, num_heads, dh)
            pars = weights["attention"][l].reshape(num_heads, 2 * dh)
            e = torch.einsum("vki,ki->vk", torch.cat([ts, tt], dim=-1), pars)
            coeffs.append(torch.where(e > 0, e, 0.2 * e))
            msgs.append(ts)
        msgs, coeffs = torch.cat(msgs), torch.cat(coeffs)
        heads = []
        for k in range(num_heads):
            x = coeffs[:, k]
            mx = segment_reduce(x.detach(), targets, V, "max")        # the max shift carries no gradient
            z = x - mx[targets]
            log_s = torch.log(segment_reduce(torch.exp(z), targets, V, "sum"))
            att = torch.exp(z - log_s[targets])
            heads.append(segment_reduce(att.unsqueeze(-1) * msgs[:, k, :], targets, V, "sum"))
        cur = act(torch.cat(heads, dim=-1))
    return cur


def sparse_gnn_film_layer(h, adjacency_lists, type_to_num_incoming_edges, num_timesteps=1, activation_function="ReLU",
                          message_aggregation_function="sum", normalize_by_num_incoming=False, *, weights: Dict):
    """gnns/gnn_film.py:58-122."""
    adj, act, V = _adj(adjacency_lists), get_activation(activation_function), h.shape[0]
    targets = torch.cat([a[:, 1] for a in adj])
    D = weights["edge_weights"][0].shape[1]
    cur = h
    for t in range(num_timesteps):
        per_type = []
        for l, a in enumerate(adj):
            m = cur[a[:, 0]] @ weights["edge_weights"][l]
            if normalize_by_num_incoming:
                m = _scale(type_to_num_incoming_edges, l, a[:, 1]) * m
            pm = (cur @ weights["film_weights"][l])[a[:, 1]]
            per_type.append(pm[:, :D] * m + pm[:, D:])
        agg = segment_reduce(act(torch.cat(per_type)), targets, V, message_aggregation_function)
        g, b = _ln(weights, t, D)
        cur = layer_norm(agg, g, b)
    return cur


def sparse_gnn_edge_mlp_layer(h, adjacency_lists, type_to_num_incoming_edges, num_timesteps=1, activation_function="ReLU",
                              message_aggregation_function="sum", normalize_by_num_incoming=False,
                              use_target_state_as_input=True, *, weights: Dict):
    """gnns/gnn_edge_mlp.py:63-122 (hidden activation of the edge MLP is ELU, :76)."""
    adj, act, V = _adj(adjacency_lists), get_activation(activation_function), h.shape[0]
    elu = get_activation("elu")
    targets = torch.cat([a[:, 1] for a in adj])
    D = weights["edge_mlps"][0][-1].shape[1]
    cur = h
    for t in range(num_timesteps):
        per_type = []
        for l, a in enumerate(adj):
            x = cur[a[:, 0]]
            if use_target_state_as_input:
                x = torch.cat([x, cur[a[:, 1]]], dim=1)
            m = mlp_apply(weights["edge_mlps"][l], x, elu)
            if normalize_by_num_incoming:
                m = _scale(type_to_num_incoming_edges, l, a[:, 1]) * m
            per_type.append(m)
        agg = segment_reduce(act(torch.cat(per_type)), targets, V, message_aggregation_function)
        g, b = _ln(weights, t, D)
        cur = layer_norm(agg, g, b)
    return cur


def sparse_rgin_layer(h, adjacency_lists, num_timesteps=1, activation_function="ReLU", message_aggregation_function="sum",
                      use_target_state_as_input=False, *, weights: Dict):
    """gnns/rgin.py:69-142 (edge MLPs / aggregation MLP are used when present in ``weights``)."""
    adj, act, V = _adj(adjacency_lists), get_activation(activation_function), h.shape[0]
    targets = torch.cat([a[:, 1] for a in adj])
    edge_mlps, aggr_mlp = weights.get("edge_mlps"), weights.get("aggr_mlp")
    cur = h
    for t in range(num_timesteps):
        per_type = []
        for l, a in enumerate(adj):
            x = cur[a[:, 0]]
            if use_target_state_as_input:
                x = torch.cat([x, cur[a[:, 1]]], dim=1)
            if edge_mlps is not None:
                x = mlp_apply(edge_mlps[l], x, act)
            per_type.append(x)
        allm = torch.cat(per_type)
        if edge_mlps is not None:
            allm = act(allm)
        new = segment_reduce(allm, targets, V, message_aggregation_function)
        if aggr_mlp is not None:
            new = mlp_apply(aggr_mlp, new, act)
        new = act(new)
        g, b = _ln(weights, t, new.shape[1])
        cur = layer_norm(new, g, b)
    return cur


def to_torch64(weights, requires_grad: bool = True):
    """numpy weight container -> the same container of float64 leaf tensors."""
    if isinstance(weights, dict):
        return {k: to_torch64(v, requires_grad) for k, v in weights.items()}
    if isinstance(weights, (list, tuple)):
        return [to_torch64(v, requires_grad) for v in weights]
    if weights is None:
        return None
    return torch.as_tensor(weights, dtype=torch.float64).clone().requires_grad_(requires_grad)


def flatten(weights, prefix: str = "") -> Dict[str, torch.Tensor]:
    """Flat {path: tensor} view of a nested weight container (stable order)."""
    out: Dict[str, torch.Tensor] = {}
    if isinstance(weights, dict):
        for k in weights:
            out.update(flatten(weights[k], prefix + str(k) + "."))
    elif isinstance(weights, (list, tuple)):
        for i, v in enumerate(weights):
            out.update(flatten(v, prefix + str(i) + "."))
    elif weights is not None:
        out[prefix[:-1]] = weights
    return out
