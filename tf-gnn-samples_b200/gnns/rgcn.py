"""sparse_rgcn_layer -- drop-in for the reference's gnns/rgcn.py:8-117 on torch CUDA tensors."""
from typing import Dict, List, Optional

import torch

from ..utils import LAYER_RGCN, get_activation, get_aggregation_function
from ._common import (check, current_stream_ptr, load_library, num_incoming_tensor, prepare, ptr_table, weight_list,
                      workspace)


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
    lib = load_library()
    out = torch.empty((plan.num_nodes, d_out), dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGCN, d_in, d_out, 0)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_rgcn_forward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(ws),
                                    cnt.data_ptr() if cnt is not None else None,
                                    act, agg, int(bool(normalize_by_num_incoming)),
                                    int(bool(use_both_source_and_target)), int(num_timesteps),
                                    out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                    current_stream_ptr(h.device)))
    return out
