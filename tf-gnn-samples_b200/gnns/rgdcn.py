"""sparse_rgdcn_layer -- drop-in for the reference's gnns/rgdcn.py:8-171 on torch CUDA tensors."""
from typing import Dict, List, Optional

import torch

from ..utils import LAYER_RGDCN, get_activation, get_aggregation_function
from ..engine import output_rows
from ._common import (RgnnError, RGNN_E_INVALID, as_f32, check, current_stream_ptr, load_library, num_incoming_tensor, prepare,
                      ptr_table, workspace)
from . import _train


def sparse_rgdcn_layer(node_embeddings: torch.Tensor,
                       adjacency_lists,
                       type_to_num_incoming_edges: Optional[torch.Tensor],
                       num_channels: int = 8,
                       channel_dim: int = 16,
                       num_timesteps: int = 1,
                       use_full_state_for_channel_weights: bool = False,
                       tie_channel_weights: bool = False,
                       activation_function: Optional[str] = "tanh",
                       message_aggregation_function: str = "sum",
                       normalize_by_num_incoming: bool = True,
                       *, weights: Dict[str, List[List[torch.Tensor]]], plan=None) -> torch.Tensor:
    """Message passing with dynamic convolutions as edge kernels (all four variants of gnns/rgdcn.py:26-53):
    h'_v[c] = act( agg_{l,(u,v)} h_u[c] . reshape(act(F_{l,c} x_v), [K, K]) / (c_{l,v} + 1e-7) ),
    x_v = h_v (use_full_state_for_channel_weights) or h_v[c].

    Same arguments as the reference plus
      weights: {"channel_weights": L x C' x [D or K, K*K]} -- the Keras kernels ``Edge_%i_Channel_%i_Weight_Computation``
               (rgdcn.py:97-104); C' = 1 when tie_channel_weights (one kernel per edge type, :96,105-107) else num_channels;
      plan:    optional GraphPlan to reuse.
    Returns float32 [V, D], D = num_channels * channel_dim.  Inference only (no gradient path in this build).
    """
    act = get_activation(activation_function)                      # rgdcn.py:92
    agg = get_aggregation_function(message_aggregation_function)   # rgdcn.py:93
    h, plan, d_in, _ = prepare(node_embeddings, adjacency_lists, plan, None)
    C, K = int(num_channels), int(channel_dim)
    if C * K != d_in:
        raise RgnnError(RGNN_E_INVALID, "sparse_rgdcn_layer: num_channels * channel_dim = %d != state dim %d" % (C * K, d_in))
    L = plan.num_edge_types
    per_type = weights["channel_weights"]
    rows = d_in if use_full_state_for_channel_weights else K
    want = 1 if tie_channel_weights else C
    if len(per_type) != L:
        raise RgnnError(RGNN_E_INVALID, "sparse_rgdcn_layer: expected %d lists under 'channel_weights', got %d" % (L, len(per_type)))
    flat = []
    for l, ws in enumerate(per_type):
        if len(ws) != want:
            raise RgnnError(RGNN_E_INVALID, "sparse_rgdcn_layer: edge type %d has %d channel kernels, expected %d" % (l, len(ws), want))
        ws = [as_f32(w, "channel_weights[%d]" % l) for w in ws]
        for w in ws:
            if tuple(w.shape) != (rows, K * K):
                raise RgnnError(RGNN_E_INVALID, "sparse_rgdcn_layer: channel kernel shape %s, expected %s" % (tuple(w.shape), (rows, K * K)))
        flat.extend(ws * C if tie_channel_weights else ws)
    if _train.requires_grad(h, flat):
        raise RgnnError(RGNN_E_INVALID, "sparse_rgdcn_layer has no gradient path in this build; call it under torch.no_grad()")
    cnt = num_incoming_tensor(type_to_num_incoming_edges, plan, normalize_by_num_incoming)
    lib = load_library()
    out = output_rows(plan, d_in, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGDCN, d_in, d_in, K)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_rgdcn_forward(plan.handle, h.data_ptr(), d_in, C, ptr_table(flat),
                                     int(bool(use_full_state_for_channel_weights)),
                                     cnt.data_ptr() if cnt is not None else None, act, agg,
                                     int(bool(normalize_by_num_incoming)), int(num_timesteps),
                                     out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(), current_stream_ptr(h.device)))
    return out
