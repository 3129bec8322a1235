"""sparse_gnn_film_layer -- drop-in for the reference's gnns/gnn_film.py:8-122 on torch CUDA tensors."""
from typing import Dict, Optional

import torch

from ..utils import LAYER_FILM, get_activation, get_aggregation_function
from ..engine import output_rows
from ._common import (RgnnError, RGNN_E_INVALID, check, current_stream_ptr, layer_norm_params, load_library, num_incoming_tensor, prepare,
                      ptr_table, weight_list, workspace)
from . import _train


def sparse_gnn_film_layer(node_embeddings: torch.Tensor,
                          adjacency_lists,
                          type_to_num_incoming_edges: Optional[torch.Tensor],
                          state_dim: Optional[int],
                          num_timesteps: int = 1,
                          activation_function: Optional[str] = "ReLU",
                          message_aggregation_function: str = "sum",
                          normalize_by_num_incoming: bool = False,
                          *, weights: Dict, plan=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """h'_v = LayerNorm( agg_{l,(u,v)} act( gamma_{l,v} * (W_l h_u)[/c] + beta_{l,v} ) ), [gamma|beta] = F_l h_v
    (gnns/gnn_film.py:85-120; the activation is inside the sum, none after).

    weights: {"edge_weights": L x [D, state_dim], "film_weights": L x [D, 2*state_dim],
              "ln_gamma"/"ln_beta": [state_dim] or one per timestep (default 1 / 0)}
    out: optional preallocated float32 [V, state_dim] result buffer (inference path only) -- sharded execution writes the
         owned rows straight into the peer-visible state buffer of the next layer (sharded.ShardedGraph.states).
    """
    act = get_activation(activation_function)
    agg = get_aggregation_function(message_aggregation_function)
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, state_dim)
    L = plan.num_edge_types
    ws = weight_list(weights, "edge_weights", L, (d_in, d_out), "sparse_gnn_film_layer")
    fw = weight_list(weights, "film_weights", L, (d_in, 2 * d_out), "sparse_gnn_film_layer")
    cnt = num_incoming_tensor(type_to_num_incoming_edges, plan, normalize_by_num_incoming)
    g, b = layer_norm_params(weights, int(num_timesteps), d_out, h.device)
    if _train.requires_grad(h, ws, fw, g, b):                         # training: differentiable composition (gnns/_train.py)
        return _train.film(h, plan, cnt, ws, fw, (g, b), act, message_aggregation_function, num_timesteps)
    lib = load_library()
    if out is None:
        out = output_rows(plan, d_out, h.device)
    elif (out.dtype != torch.float32 or tuple(out.shape) != (plan.num_nodes, d_out) or not out.is_contiguous()
          or out.device != h.device or out.data_ptr() == h.data_ptr()):
        raise RgnnError(RGNN_E_INVALID, "out must be a contiguous float32 [%d, %d] tensor on %s that does not alias the input"
                        % (plan.num_nodes, d_out, h.device))
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_FILM, d_in, d_out, 0)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_film_forward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(ws), ptr_table(fw),
                                    cnt.data_ptr() if cnt is not None else None, g.data_ptr(), b.data_ptr(),
                                    act, agg, int(bool(normalize_by_num_incoming)), int(num_timesteps),
                                    out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                    current_stream_ptr(h.device)))
    return out
