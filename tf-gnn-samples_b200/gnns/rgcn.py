"""sparse_rgcn_layer -- drop-in for the reference's gnns/rgcn.py:8-117 on torch CUDA tensors."""
from typing import Dict, List, Optional

import torch

from ..utils import AGG_MAX, LAYER_RGCN, LAYER_RGCN_BACKWARD, get_activation, get_aggregation_function
from ..engine import output_rows
from ._common import (check, current_stream_ptr, load_library, num_incoming_tensor, prepare, ptr_table, weight_list,
                      workspace)
from . import _train


def _forward_raw(h, plan, cnt, ws, d_in, d_out, act, agg, normalize, both, num_timesteps):
    lib = load_library()
    out = output_rows(plan, d_out, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGCN, d_in, d_out, 0)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_rgcn_forward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(ws),
                                    cnt.data_ptr() if cnt is not None else None,
                                    act, agg, int(bool(normalize)), int(bool(both)), int(num_timesteps),
                                    out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                    current_stream_ptr(h.device)))
    return out


class _RGCNStep(torch.autograd.Function):
    """One timestep of sparse_rgcn_layer with gradients for the node states and the per-type kernels
    (rgnn_rgcn_backward).  The reference differentiates gnns/rgcn.py:84-114 with TF autodiff
    (models/sparse_graph_model.py:253-260); this is the same gradient, computed by dedicated kernels."""

    @staticmethod
    def forward(ctx, h, plan, cnt, act, agg, normalize, *ws):
        d_in, d_out = int(h.shape[1]), int(ws[0].shape[1])
        out = _forward_raw(h, plan, cnt, list(ws), d_in, d_out, act, agg, normalize, False, 1)
        ctx.plan, ctx.cnt, ctx.act, ctx.agg, ctx.normalize = plan, cnt, act, agg, normalize
        ctx.save_for_backward(h, out, *ws)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        h, out, *ws = ctx.saved_tensors
        plan, cnt = ctx.plan, ctx.cnt
        d_in, d_out = int(h.shape[1]), int(ws[0].shape[1])
        g = grad_out.contiguous().float()
        need_h = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[6:])
        grad_h = torch.empty_like(h) if need_h else None
        grad_ws = [torch.empty_like(w) for w in ws] if need_w else None
        lib = load_library()
        with torch.cuda.device(h.device):
            nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGCN_BACKWARD, d_in, d_out, 0)
            ws_buf = workspace(h.device, nbytes)
            check(lib.rgnn_rgcn_backward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(list(ws)),
                                         cnt.data_ptr() if cnt is not None else None, ctx.act, ctx.agg,
                                         int(bool(ctx.normalize)), out.data_ptr(), g.data_ptr(),
                                         grad_h.data_ptr() if need_h else None,
                                         ptr_table(grad_ws, weights=False) if need_w else None,
                                         ws_buf.data_ptr(), ws_buf.numel(), current_stream_ptr(h.device)))
        return (grad_h, None, None, None, None, None) + (tuple(grad_ws) if need_w else tuple(None for _ in ws))


def sparse_rgcn_layer(node_embeddings: torch.Tensor,
                      adjacency_lists,
                      type_to_num_incoming_edges: Optional[torch.Tensor],
                      state_dim: Optional[int],
                      num_timesteps: int = 1,
                      activation_function: Optional[str] = "tanh",
                      message_aggregation_function: str = "sum",
                      normalize_by_num_incoming: bool = True,
                      use_both_source_and_target: bool = False,
                      *, weights: Dict[str, List[torch.Tensor]], plan=None) -> torch.Tensor:
    """h'_v = act( agg_{l, (u,v) in A_l} (W_l h_u [| h_v]) / (c_{l,v} + 1e-7) ), repeated num_timesteps.

    Same arguments as the reference (gnns/rgcn.py:8-17) plus:
      weights: {"edge_weights": L x [D*(1+use_both_source_and_target), state_dim]} -- the Keras kernels
               the reference creates as ``Edge_%i_Weight`` (rgcn.py:69-75);
      plan:    optional GraphPlan to reuse across layers (else built from adjacency_lists).
    Returns float32 [V, state_dim] on the input's device.
    """
    act = get_activation(activation_function)                      # rgcn.py:72 (ValueError on unknown)
    agg = get_aggregation_function(message_aggregation_function)   # rgcn.py:73
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, state_dim)
    L = plan.num_edge_types
    k_rows = d_in * (2 if use_both_source_and_target else 1)
    ws = weight_list(weights, "edge_weights", L, (k_rows, d_out), "sparse_rgcn_layer")
    cnt = num_incoming_tensor(type_to_num_incoming_edges, plan, normalize_by_num_incoming)
    needs_grad = torch.is_grad_enabled() and (h.requires_grad or any(w.requires_grad for w in ws))
    if not needs_grad:
        return _forward_raw(h, plan, cnt, ws, d_in, d_out, act, agg, normalize_by_num_incoming,
                            use_both_source_and_target, num_timesteps)
    # training path: one differentiable step per timestep (autograd chains them)
    if use_both_source_and_target or agg == AGG_MAX:
        # settings outside the fused backward kernels: differentiable composition of the engine's building blocks
        return _train.rgcn(h, plan, cnt, ws, act, message_aggregation_function, bool(use_both_source_and_target), num_timesteps)
    cur = h
    for _ in range(int(num_timesteps)):
        cur = _RGCNStep.apply(cur, plan, cnt, act, agg, bool(normalize_by_num_incoming), *ws)
    return cur


def rgcn_layer_stack(node_embeddings: torch.Tensor, adjacency_lists, type_to_num_incoming_edges,
                     layer_weights: List[Dict[str, List[torch.Tensor]]],
                     activation_function: Optional[str] = "ReLU", message_aggregation_function: str = "sum",
                     normalize_by_num_incoming: bool = True, *, plan=None) -> torch.Tensor:
    """graph_num_layers x sparse_rgcn_layer in ONE library call (rgnn_rgcn_stack_forward): the GNN loop of
    Sparse_Graph_Model.__build_graph_propagation_model (models/sparse_graph_model.py:176-191) for RGCN_Model,
    without the scaffold's dropout / residual / inter-layer Dense.  Same result as calling sparse_rgcn_layer
    once per entry of ``layer_weights``; saves the per-layer host overhead."""
    act = get_activation(activation_function)
    agg = get_aggregation_function(message_aggregation_function)
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, None)
    L = plan.num_edge_types
    flat = []
    for w in layer_weights:
        flat.extend(weight_list(w, "edge_weights", L, (d_in, d_in), "rgcn_layer_stack"))
    cnt = num_incoming_tensor(type_to_num_incoming_edges, plan, normalize_by_num_incoming)
    lib = load_library()
    out = output_rows(plan, d_in, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGCN, d_in, d_in, 0) + 2 * (plan.num_nodes * d_in * 4 + 256)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_rgcn_stack_forward(plan.handle, h.data_ptr(), d_in, len(layer_weights), ptr_table(flat),
                                          cnt.data_ptr() if cnt is not None else None, act, agg,
                                          int(bool(normalize_by_num_incoming)), out.data_ptr(), ws_buf.data_ptr(),
                                          ws_buf.numel(), current_stream_ptr(h.device)))
    return out
