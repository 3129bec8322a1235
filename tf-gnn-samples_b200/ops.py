"""Building blocks of the hot path exported on their own (include/rgnn.h 'building blocks'):
the tensor-core Dense, the segment aggregation of utils.get_aggregation_function, and layer norm."""
from typing import Optional

import torch

from .engine import GraphPlan, RgnnError, RGNN_E_INVALID, as_f32, check, current_stream_ptr, load_library
from .utils import get_activation, get_aggregation_function


def dense(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor] = None,
          activation: Optional[str] = None) -> torch.Tensor:
    """act(x @ kernel + bias): tf.keras.layers.Dense with the Keras [in, out] kernel,
    fp32-accurate on the tensor cores (tcgen05 3xTF32)."""
    x, kernel = as_f32(x, "x"), as_f32(kernel, "kernel")
    if x.dim() != 2 or kernel.dim() != 2 or x.shape[1] != kernel.shape[0]:
        raise RgnnError(RGNN_E_INVALID, "dense: shapes %s x %s do not contract" % (tuple(x.shape), tuple(kernel.shape)))
    b = as_f32(bias, "bias") if bias is not None else None
    out = torch.empty((x.shape[0], kernel.shape[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load_library().rgnn_dense_forward(x.data_ptr(), x.shape[0], x.shape[1], kernel.data_ptr(), kernel.shape[1],
                                                b.data_ptr() if b is not None else None, get_activation(activation),
                                                out.data_ptr(), current_stream_ptr(x.device)))
    return out


def segment_aggregate(plan: GraphPlan, data: torch.Tensor, aggregation: str = "sum") -> torch.Tensor:
    """tf.unsorted_segment_<agg>(data, message_targets, num_nodes) for `data` [M, d] whose rows are in the
    type-major concatenation order of the adjacency lists the plan was built from (gnns/rgcn.py:108-112)."""
    data = as_f32(data, "data")
    if data.dim() != 2 or data.shape[0] != plan.num_edges:
        raise RgnnError(RGNN_E_INVALID, "segment_aggregate: data must be [M=%d, d], got %s" % (plan.num_edges, tuple(data.shape)))
    out = torch.empty((plan.num_nodes, data.shape[1]), dtype=torch.float32, device=data.device)
    with torch.cuda.device(data.device):
        check(load_library().rgnn_segment_aggregate(plan.handle, data.data_ptr(), data.shape[1],
                                                    get_aggregation_function(aggregation), out.data_ptr(),
                                                    current_stream_ptr(data.device)))
    return out


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """tf.contrib.layers.layer_norm defaults: last axis, biased variance, eps 1e-12."""
    x, gamma, beta = as_f32(x, "x"), as_f32(gamma, "gamma"), as_f32(beta, "beta")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(load_library().rgnn_layer_norm(x.data_ptr(), x.shape[0], x.shape[1], gamma.data_ptr(), beta.data_ptr(),
                                             out.data_ptr(), current_stream_ptr(x.device)))
    return out
