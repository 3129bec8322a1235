"""Analytic gradients of the RGCN layer in numpy (TEST INFRASTRUCTURE -- see oracle/__init__.py; derivatives of the PINNED
forward oracle.ref_layers.sparse_rgcn_layer, checked by finite differences -- TF autodiff itself cannot run here).

The reference gets its gradients from TensorFlow autodiff of gnns/rgcn.py:84-114
(models/sparse_graph_model.py:253-260 calls tf.gradients on the loss).  This restates what autodiff
produces for one timestep of that op sequence; tests/test_oracle_grads.py pins it against central finite
differences of oracle.ref_layers.sparse_rgcn_layer in float64.
"""
from typing import Dict, List, Optional

import numpy as np

from . import ref_layers as R


def activation_grad(name: Optional[str], pre: np.ndarray) -> np.ndarray:
    """d act(x) / dx at the pre-activation x (utils/utils.py:36-58 activations)."""
    if name is None or name.lower() == "linear":
        return np.ones_like(pre)
    n = name.lower()
    if n == "tanh":
        return 1.0 - np.tanh(pre) ** 2
    if n == "relu":
        return (pre > 0).astype(pre.dtype)
    if n == "leaky_relu":
        return np.where(pre > 0, 1.0, 0.2).astype(pre.dtype)
    if n == "elu":
        return np.where(pre > 0, 1.0, np.exp(np.minimum(pre, 0))).astype(pre.dtype)
    if n == "selu":
        scale, alpha = 1.0507009873554805, 1.6732632423543772
        return (scale * np.where(pre > 0, 1.0, alpha * np.exp(np.minimum(pre, 0)))).astype(pre.dtype)
    if n == "gelu":
        from scipy.special import erf
        return (0.5 * (1.0 + erf(pre / np.sqrt(2.0))) + pre * np.exp(-0.5 * pre * pre) / np.sqrt(2.0 * np.pi)).astype(pre.dtype)
    raise ValueError("Unknown activation function '%s'!" % name)


def rgcn_layer_grads(node_embeddings, adjacency_lists, type_to_num_incoming_edges, grad_out, activation_function="tanh",
                     message_aggregation_function="sum", normalize_by_num_incoming=True, *, weights: Dict,
                     dtype=np.float64):
    """Gradients of sum(out * grad_out) for ONE timestep of sparse_rgcn_layer (source-only messages)
    w.r.t. the node states and the per-type kernels.  Returns (d_h [V, Din], [d_W_l [Din, D]])."""
    h = np.asarray(node_embeddings, dtype)
    adj = [np.asarray(a).reshape(-1, 2).astype(np.int64) for a in adjacency_lists]
    V = h.shape[0]
    ws = [np.asarray(w, dtype) for w in weights["edge_weights"]]
    D = ws[0].shape[1]
    g = np.asarray(grad_out, dtype)
    cnt = np.asarray(type_to_num_incoming_edges, dtype) if type_to_num_incoming_edges is not None else None
    tgt_all = np.concatenate([a[:, 1] for a in adj])
    # forward pieces needed by the backward (gnns/rgcn.py:88-110)
    scales, msgs = [], []
    for l, a in enumerate(adj):
        s = (1.0 / (cnt[l][a[:, 1]] + R.SMALL_NUMBER)) if normalize_by_num_incoming else np.ones(a.shape[0], dtype)
        scales.append(s)
        msgs.append(s[:, None] * (h[a[:, 0]] @ ws[l]))
    allm = np.concatenate(msgs, axis=0) if adj else np.zeros((0, D), dtype)
    if message_aggregation_function not in ("sum", "mean", "sqrt_n"):
        raise NotImplementedError("gradient of '%s' aggregation is not restated" % message_aggregation_function)
    agg = R.unsorted_segment_sum(allm, tgt_all, V)
    n = np.maximum(np.bincount(tgt_all, minlength=V).astype(dtype), 1.0)
    div = {"sum": np.ones(V, dtype), "mean": n, "sqrt_n": np.sqrt(n)}[message_aggregation_function]
    pre = agg / div[:, None]
    d_pre = g * activation_grad(activation_function, pre)
    d_agg = d_pre / div[:, None]
    d_h = np.zeros_like(h)
    d_ws: List[np.ndarray] = []
    for l, a in enumerate(adj):
        src, tgt = a[:, 0], a[:, 1]
        d_msg = scales[l][:, None] * d_agg[tgt]                 # [E, D]
        d_ws.append(h[src].T @ d_msg)                           # [Din, D]
        np.add.at(d_h, src, d_msg @ ws[l].T)
    return d_h, d_ws
