"""sparse_rgin_layer -- drop-in for the reference's gnns/rgin.py:7-142 on torch CUDA tensors."""
from typing import Dict, Optional

import torch

from ..utils import LAYER_RGIN, get_activation, get_aggregation_function
from ..engine import output_rows
from ._common import (RgnnError, RGNN_E_INVALID, as_f32, check, current_stream_ptr, int32_array, layer_norm_params,
                      load_library, mlp_tables, prepare, ptr_table, workspace)
from . import _train


def sparse_rgin_layer(node_embeddings: torch.Tensor,
                      adjacency_lists,
                      state_dim: Optional[int],
                      num_timesteps: int = 1,
                      activation_function: Optional[str] = "ReLU",
                      message_aggregation_function: str = "sum",
                      use_target_state_as_input: bool = False,
                      num_edge_MLP_hidden_layers: Optional[int] = 1,
                      num_aggr_MLP_hidden_layers: Optional[int] = None,
                      *, weights: Dict, plan=None) -> torch.Tensor:
    """h'_v = LayerNorm( act( MLP_aggr( agg_{l,(u,v)} act(MLP_l(h_u [|| h_v])) ) ) )  (gnns/rgin.py:103-140).

    ``num_edge_MLP_hidden_layers=None`` -> raw source states are the messages (no activation);
    ``num_aggr_MLP_hidden_layers=None`` -> no aggregation MLP.  MLP hidden activation = activation_function.
    weights: {"edge_mlps": L x [kernels...] (``Edge_%i_MLP``), "aggr_mlp": [kernels...] (``Aggregation_MLP``),
              "ln_gamma"/"ln_beta"}
    """
    act = get_activation(activation_function)
    agg = get_aggregation_function(message_aggregation_function)
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, state_dim)
    L = plan.num_edge_types
    edge_ptrs, edge_dims, n_edge_hidden, nl_edge = None, None, -1, 0
    if num_edge_MLP_hidden_layers is not None:                                   # rgin.py:86-89
        mlps = weights["edge_mlps"]
        if len(mlps) != L:
            raise RgnnError(RGNN_E_INVALID, "sparse_rgin_layer: expected %d edge MLPs, got %d" % (L, len(mlps)))
        flat, dims, nl_edge = mlp_tables(mlps, "edge_mlps")
        if nl_edge != int(num_edge_MLP_hidden_layers) + 1:
            raise RgnnError(RGNN_E_INVALID, "sparse_rgin_layer: num_edge_MLP_hidden_layers=%d needs %d kernels per "
                            "type, got %d" % (num_edge_MLP_hidden_layers, num_edge_MLP_hidden_layers + 1, nl_edge))
        edge_keep, edge_ptrs, edge_dims, n_edge_hidden = flat, ptr_table(flat), int32_array(dims), int(num_edge_MLP_hidden_layers)
    aggr_ptrs, aggr_dims, n_aggr_hidden, nl_aggr = None, None, -1, 0
    if num_aggr_MLP_hidden_layers is not None:                                   # rgin.py:78-84
        ks = [as_f32(k, "aggr_mlp") for k in weights["aggr_mlp"]]
        nl_aggr = len(ks)
        if nl_aggr != int(num_aggr_MLP_hidden_layers) + 1:
            raise RgnnError(RGNN_E_INVALID, "sparse_rgin_layer: num_aggr_MLP_hidden_layers=%d needs %d kernels, got %d"
                            % (num_aggr_MLP_hidden_layers, num_aggr_MLP_hidden_layers + 1, nl_aggr))
        adims = [int(ks[0].shape[0])] + [int(k.shape[1]) for k in ks]
        aggr_keep, aggr_ptrs, aggr_dims, n_aggr_hidden = ks, ptr_table(ks), int32_array(adims), int(num_aggr_MLP_hidden_layers)
    g, b = layer_norm_params(weights, int(num_timesteps), d_out, h.device)
    if _train.requires_grad(h, weights.get("edge_mlps") if edge_ptrs is not None else None,
                            weights.get("aggr_mlp") if aggr_ptrs is not None else None, g, b):   # training (gnns/_train.py)
        per_type = [edge_keep[l * nl_edge:(l + 1) * nl_edge] for l in range(L)] if edge_ptrs is not None else None
        return _train.rgin(h, plan, per_type, aggr_keep if aggr_ptrs is not None else None, (g, b), act,
                           message_aggregation_function, bool(use_target_state_as_input), num_timesteps)
    lib = load_library()
    out = output_rows(plan, d_out, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGIN, d_in, d_out, max(nl_edge, nl_aggr))
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_rgin_forward(plan.handle, h.data_ptr(), d_in, d_out, edge_ptrs, edge_dims, n_edge_hidden,
                                    aggr_ptrs, aggr_dims, n_aggr_hidden, g.data_ptr(), b.data_ptr(),
                                    act, agg, int(bool(use_target_state_as_input)), int(num_timesteps),
                                    out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                    current_stream_ptr(h.device)))
    return out
