"""sparse_rgat_layer -- drop-in for the reference's gnns/rgat.py:9-141 on torch CUDA tensors."""
from typing import Dict, Optional

import torch

from ..utils import LAYER_RGAT, get_activation
from ..engine import output_rows
from ._common import (check, current_stream_ptr, load_library, prepare, ptr_table, weight_list, workspace)
from . import _train


def sparse_rgat_layer(node_embeddings: torch.Tensor,
                      adjacency_lists,
                      state_dim: Optional[int],
                      num_heads: int = 4,
                      num_timesteps: int = 1,
                      activation_function: Optional[str] = "tanh",
                      *, weights: Dict, plan=None) -> torch.Tensor:
    """Relational GAT: per type T_l = H W_l; per edge and head k a LeakyReLU(0.2) logit from the
    transformed source and target states; softmax over ALL incoming messages of a node (all types);
    weighted sum per head; activation (gnns/rgat.py:83-139).

    weights: {"edge_weights": L x [D, state_dim], "attention": L x [2 * state_dim]}
             (``Edge_%i_Attention_Parameters``, rgat.py:74-76; head k uses the slice
             [k*2d, (k+1)*2d): first d entries for the source, next d for the target, rgat.py:110-111).
    """
    act = get_activation(activation_function)
    h, plan, d_in, d_out = prepare(node_embeddings, adjacency_lists, plan, state_dim)
    L = plan.num_edge_types
    ws = weight_list(weights, "edge_weights", L, (d_in, d_out), "sparse_rgat_layer")
    att = weight_list(weights, "attention", L, (2 * d_out,), "sparse_rgat_layer")
    if d_out % int(num_heads):
        raise RgnnError(RGNN_E_INVALID, "sparse_rgat_layer: state_dim %d is not divisible by num_heads %d" % (d_out, num_heads))
    if _train.requires_grad(h, ws, att):                              # training: differentiable composition (gnns/_train.py)
        return _train.rgat(h, plan, ws, att, int(num_heads), act, num_timesteps)
    lib = load_library()
    out = output_rows(plan, d_out, h.device)
    with torch.cuda.device(h.device):
        nbytes = lib.rgnn_workspace_bytes(plan.handle, LAYER_RGAT, d_in, d_out, 0)
        ws_buf = workspace(h.device, nbytes)
        check(lib.rgnn_rgat_forward(plan.handle, h.data_ptr(), d_in, d_out, ptr_table(ws), ptr_table(att),
                                    int(num_heads), act, int(num_timesteps),
                                    out.data_ptr(), ws_buf.data_ptr(), ws_buf.numel(),
                                    current_stream_ptr(h.device)))
    return out
