"""Argument plumbing shared by the layer functions."""
from typing import Dict, List, Optional, Sequence

import torch

from ..engine import (GraphPlan, RgnnError, RGNN_E_INVALID, as_f32, c_int32, check, current_stream_ptr,
                      load_library, ptr_table, resolve_plan, workspace)


def prepare(node_embeddings: torch.Tensor, adjacency_lists, plan: Optional[GraphPlan], state_dim: Optional[int]):
    h = as_f32(node_embeddings, "node_embeddings")
    if h.dim() != 2:
        raise RgnnError(RGNN_E_INVALID, "node_embeddings must be [V, D], got shape %s" % (tuple(h.shape),))
    plan = resolve_plan(h, adjacency_lists, plan)
    if plan.num_nodes != h.shape[0]:
        raise RgnnError(RGNN_E_INVALID, "plan was built for %d nodes, node_embeddings has %d rows"
                        % (plan.num_nodes, h.shape[0]))
    d_in = int(h.shape[1])
    d_out = d_in if state_dim is None else int(state_dim)   # gnns/rgcn.py:68-69
    return h, plan, d_in, d_out


def weight_list(weights: Dict, key: str, count: int, shape, what: str) -> List[torch.Tensor]:
    ws = weights[key]
    if len(ws) != count:
        raise RgnnError(RGNN_E_INVALID, "%s: expected %d tensors under '%s', got %d" % (what, count, key, len(ws)))
    out = []
    for i, w in enumerate(ws):
        w = as_f32(w, "%s[%d]" % (key, i))
        if shape is not None and tuple(w.shape) != tuple(shape):
            raise RgnnError(RGNN_E_INVALID, "%s: %s[%d] has shape %s, expected %s"
                            % (what, key, i, tuple(w.shape), tuple(shape)))
        out.append(w)
    return out


def num_incoming_tensor(type_to_num_incoming_edges, plan: GraphPlan, needed: bool) -> Optional[torch.Tensor]:
    if not needed:
        return None
    if type_to_num_incoming_edges is None:
        raise RgnnError(RGNN_E_INVALID, "normalize_by_num_incoming=True needs type_to_num_incoming_edges")
    c = type_to_num_incoming_edges
    if not isinstance(c, torch.Tensor):
        c = torch.as_tensor(c)
    c = c.to(plan.device, dtype=torch.float32).contiguous()    # fp32 placeholder: tasks/sparse_graph_task.py:145
    if tuple(c.shape) != (plan.num_edge_types, plan.num_nodes):
        raise RgnnError(RGNN_E_INVALID, "type_to_num_incoming_edges must be [L=%d, V=%d], got %s"
                        % (plan.num_edge_types, plan.num_nodes, tuple(c.shape)))
    return c


def layer_norm_params(weights: Dict, num_timesteps: int, dim: int, device) -> (torch.Tensor, torch.Tensor):
    """[T, D] gamma / beta.  The reference opens a fresh LayerNorm scope per timestep
    (gnn_film.py:120), so a list with one [D] vector per timestep is accepted; a single [D] vector is
    shared; missing -> the tf.contrib.layers.layer_norm initial values gamma=1, beta=0."""
    def one(key, fill):
        v = weights.get(key)
        if v is None:
            return torch.full((num_timesteps, dim), fill, dtype=torch.float32, device=device)
        if isinstance(v, (list, tuple)):
            v = torch.stack([as_f32(x, key) for x in v], dim=0)
        v = as_f32(v, key)
        if v.dim() == 1:
            v = v.unsqueeze(0).expand(num_timesteps, dim)
        if tuple(v.shape) != (num_timesteps, dim):
            raise RgnnError(RGNN_E_INVALID, "%s must be [%d] or [%d, %d], got %s"
                            % (key, dim, num_timesteps, dim, tuple(v.shape)))
        return v.contiguous()
    return one("ln_gamma", 1.0), one("ln_beta", 0.0)


def mlp_tables(mlps: Sequence[Sequence[torch.Tensor]], what: str):
    """Type-major pointer table + dims of per-type MLPs (utils/utils.py:77-126: bias-free Dense stack)."""
    nl = len(mlps[0])
    flat, dims = [], None
    for l, ks in enumerate(mlps):
        if len(ks) != nl:
            raise RgnnError(RGNN_E_INVALID, "%s: MLP of edge type %d has %d layers, expected %d" % (what, l, len(ks), nl))
        ks = [as_f32(k, "%s[%d]" % (what, l)) for k in ks]
        d = [int(ks[0].shape[0])] + [int(k.shape[1]) for k in ks]
        for j in range(1, nl):
            if int(ks[j].shape[0]) != d[j]:
                raise RgnnError(RGNN_E_INVALID, "%s: layer %d of type %d has %d input rows, expected %d"
                                % (what, j, l, int(ks[j].shape[0]), d[j]))
        if dims is None:
            dims = d
        elif dims != d:
            raise RgnnError(RGNN_E_INVALID, "%s: MLP shapes differ between edge types" % what)
        flat.extend(ks)
    return flat, dims, nl


def int32_array(values: Sequence[int]):
    return (c_int32 * max(len(values), 1))(*[int(v) for v in values])


__all__ = ["prepare", "weight_list", "num_incoming_tensor", "layer_norm_params", "mlp_tables", "int32_array",
           "GraphPlan", "RgnnError", "RGNN_E_INVALID", "as_f32", "check", "current_stream_ptr", "load_library",
           "ptr_table", "workspace", "torch"]
