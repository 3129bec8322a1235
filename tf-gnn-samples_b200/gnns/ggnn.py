"""sparse_ggnn_layer -- drop-in for the reference's gnns/ggnn.py:8-95 on torch CUDA tensors."""
from typing import Dict, Optional

import torch

from ..utils import LAYER_GGNN, CELL_GRU, get_aggregation_function, get_gated_unit
from ..engine import note_weights
from ..engine import output_rows
from ._common import (RgnnError, RGNN_E_INVALID, as_f32, check, current_stream_ptr, load_library, prepare, ptr_table,
                      weight_list, workspace)
from . import _train


def sparse_ggnn_layer(node_embeddings: torch.Tensor,
                      adjacency_lists,
                      state_dim: Optional[int],
                      num_timesteps: int = 1,
                      gated_unit_type: str = "gru",
                      activation_function: str = "tanh",
                      message_aggregation_function: str = "sum",
                      *, weights: Dict, plan=None) -> torch.Tensor:
    """h' = Cell(inputs = agg_{l,(u,v)} W_l h_u, state = h) per timestep (gnns/ggnn.py:71-93).

    weights: {"edge_weights": L x [D, D],
              "cell": {"kernel": [D, 3D] (GRU, gates z|r|h) or [D, D] (RNN), "recurrent_kernel": same, "bias": [3D]/[D]}}
    Keras TF-1.13 cell defaults are reproduced: GRU recurrent_activation = hard_sigmoid,
    reset_after = False.
    """
    agg = get_aggregation_function(message_aggregation_function)               # ggnn.py:55
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, state_dim)
    cell_kind, act = get_gated_unit(d_out, gated_unit_type, activation_function)   # ggnn.py:56
    L = plan.num_edge_types
    ws = weight_list(weights, "edge_weights", L, (d_in, d_out), "sparse_ggnn_layer")
    cell = weights["cell"]
    gates = 3 if cell_kind == CELL_GRU else 1
    kernel = as_f32(cell["kernel"], "cell.kernel")
    rec = as_f32(cell["recurrent_kernel"], "cell.recurrent_kernel")
    bias = as_f32(cell["bias"], "cell.bias")
    if tuple(kernel.shape) != (d_out, gates * d_out) or tuple(rec.shape) != (d_out, gates * d_out) \
            or tuple(bias.shape) != (gates * d_out,):
        raise RgnnError(RGNN_E_INVALID, "sparse_ggnn_layer: cell weights must be [%d,%d], [%d,%d], [%d]; got %s %s %s"
                        % (d_out, gates * d_out, d_out, gates * d_out, gates * d_out,
                           tuple(kernel.shape), tuple(rec.shape), tuple(bias.shape)))
    if _train.requires_grad(h, ws, kernel, rec, bias):               # training: differentiable composition (gnns/_train.py)
        return _train.ggnn(h, plan, ws, {"kernel": kernel, "recurrent_kernel": rec, "bias": bias},
                           "gru" if cell_kind == CELL_GRU else "rnn", act, message_aggregation_function, num_timesteps)
    note_weights([kernel, rec, bias])
    lib = load_library()
    out = output_rows(plan, d_out, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_GGNN, d_in, d_out, 0)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_ggnn_forward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(ws), kernel.data_ptr(),
                                    rec.data_ptr(), bias.data_ptr(), cell_kind, act, agg, int(num_timesteps),
                                    out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                    current_stream_ptr(h.device)))
    return out
