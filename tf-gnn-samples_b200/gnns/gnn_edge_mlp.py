"""sparse_gnn_edge_mlp_layer -- drop-in for the reference's gnns/gnn_edge_mlp.py:7-122 on torch CUDA tensors."""
from typing import Dict, Optional

import torch

from ..utils import LAYER_EDGE_MLP, get_activation, get_aggregation_function
from ..engine import output_rows
from ._common import (RgnnError, RGNN_E_INVALID, check, current_stream_ptr, int32_array, layer_norm_params,
                      load_library, mlp_tables, num_incoming_tensor, prepare, ptr_table, workspace)
from . import _train


def sparse_gnn_edge_mlp_layer(node_embeddings: torch.Tensor,
                              adjacency_lists,
                              type_to_num_incoming_edges: Optional[torch.Tensor],
                              state_dim: Optional[int],
                              num_timesteps: int = 1,
                              activation_function: Optional[str] = "ReLU",
                              message_aggregation_function: str = "sum",
                              normalize_by_num_incoming: bool = False,
                              use_target_state_as_input: bool = True,
                              num_edge_hidden_layers: int = 1,
                              *, weights: Dict, plan=None) -> torch.Tensor:
    """h'_v = LayerNorm( agg_{l,(u,v)} act( [1/c] MLP_l(h_u || h_v) ) )  (gnns/gnn_edge_mlp.py:84-119).

    The edge MLP's hidden activation is ELU whatever ``activation_function`` says (gnn_edge_mlp.py:76).
    weights: {"edge_mlps": L x [kernel_0 [D*(1+use_target), S], ..., kernel_n [S, S]] with
              n = num_edge_hidden_layers (``Edge_%i_MLP``), "ln_gamma"/"ln_beta"}
    """
    act = get_activation(activation_function)
    agg = get_aggregation_function(message_aggregation_function)
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, state_dim)
    L = plan.num_edge_types
    mlps = weights["edge_mlps"]
    if len(mlps) != L:
        raise RgnnError(RGNN_E_INVALID, "sparse_gnn_edge_mlp_layer: expected %d edge MLPs, got %d" % (L, len(mlps)))
    flat, dims, nl = mlp_tables(mlps, "edge_mlps")
    if nl != int(num_edge_hidden_layers) + 1:
        raise RgnnError(RGNN_E_INVALID, "sparse_gnn_edge_mlp_layer: num_edge_hidden_layers=%d needs %d kernels per "
                        "type, got %d" % (num_edge_hidden_layers, num_edge_hidden_layers + 1, nl))
    cnt = num_incoming_tensor(type_to_num_incoming_edges, plan, normalize_by_num_incoming)
    g, b = layer_norm_params(weights, int(num_timesteps), d_out, h.device)
    if _train.requires_grad(h, flat, g, b):                           # training: differentiable composition (gnns/_train.py)
        per_type = [flat[l * nl:(l + 1) * nl] for l in range(L)]
        return _train.edge_mlp(h, plan, cnt, per_type, (g, b), act, message_aggregation_function,
                               bool(use_target_state_as_input), num_timesteps)
    lib = load_library()
    out = output_rows(plan, d_out, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_EDGE_MLP, d_in, d_out, nl)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_edge_mlp_forward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(flat), int32_array(dims),
                                        int(num_edge_hidden_layers),
                                        cnt.data_ptr() if cnt is not None else None, g.data_ptr(), b.data_ptr(),
                                        act, agg, int(bool(normalize_by_num_incoming)),
                                        int(bool(use_target_state_as_input)), int(num_timesteps),
                                        out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                        current_stream_ptr(h.device)))
    return out
