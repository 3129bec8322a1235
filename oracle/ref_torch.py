"""torch-CPU float32 restatement of the reference layers in the reference's op order
(TEST / BASELINE INFRASTRUCTURE -- see oracle/__init__.py).  PINNED: tests/test_cpu_baseline_pin.py compares it with the
output of the reference's own gnns/rgcn.py at BASELINE config 2 (tests/golden/ref_config2_rgcn_ppi.npz) and with the numpy oracle.

This is the timed stand-in for "the reference TF1 CPU path" (BASELINE.md 3): per edge type
index_select (tf.nn.embedding_lookup, gnns/rgcn.py:88) -> [E_l, D] @ [D, D] (Dense, :98) ->
* 1/(c + 1e-7) (:100-104) -> cat (:108) -> index_add_ (tf.unsorted_segment_sum, :110) -> activation
(:114), multi-threaded through torch's intra-op pool like TF's Eigen pool.
"""
from typing import Dict, List, Optional

import torch

SMALL_NUMBER = 1e-7

_ACTS = {
    None: lambda x: x, "linear": lambda x: x, "tanh": torch.tanh, "relu": torch.relu,
    "leaky_relu": lambda x: torch.nn.functional.leaky_relu(x, 0.2), "elu": torch.nn.functional.elu,
    "selu": torch.selu, "gelu": lambda x: torch.nn.functional.gelu(x),
}


def sparse_rgcn_layer(node_embeddings: torch.Tensor, adjacency_lists: List[torch.Tensor],
                      type_to_num_incoming_edges: torch.Tensor, state_dim: Optional[int], num_timesteps: int = 1,
                      activation_function: Optional[str] = "tanh", message_aggregation_function: str = "sum",
                      normalize_by_num_incoming: bool = True, *, weights: Dict) -> torch.Tensor:
    """gnns/rgcn.py:67-117 (sum aggregation, source-only messages: the configuration RGCN_Model uses)."""
    assert message_aggregation_function == "sum"
    act = _ACTS[activation_function.lower() if activation_function else None]
    num_nodes = node_embeddings.shape[0]
    message_targets = torch.cat([a[:, 1] for a in adjacency_lists], dim=0)                 # :78
    cur = node_embeddings
    for _ in range(num_timesteps):                                                       # :81
        per_type = []
        for l, a in enumerate(adjacency_lists):                                          # :84
            src, tgt = a[:, 0], a[:, 1]
            s = cur.index_select(0, src)                                                 # :88
            msg = s @ weights["edge_weights"][l]                                         # :98
            if normalize_by_num_incoming:                                                # :100-104
                c = type_to_num_incoming_edges[l].index_select(0, tgt)
                msg = (1.0 / (c + SMALL_NUMBER)).unsqueeze(-1) * msg
            per_type.append(msg)
        msgs = torch.cat(per_type, dim=0)                                                # :108
        agg = torch.zeros((num_nodes, msgs.shape[1]), dtype=msgs.dtype).index_add_(0, message_targets, msgs)  # :110
        cur = act(agg)                                                                   # :114
    return cur


def rgcn_stack(node_embeddings, adjacency_lists, type_to_num_incoming_edges, layer_weights: List[Dict],
               activation_function="ReLU") -> torch.Tensor:
    """graph_num_layers x sparse_rgcn_layer: the GNN part of models/sparse_graph_model.py:176-191."""
    cur = node_embeddings
    for w in layer_weights:
        cur = sparse_rgcn_layer(cur, adjacency_lists, type_to_num_incoming_edges, cur.shape[1],
                                activation_function=activation_function, weights=w)
    return cur
