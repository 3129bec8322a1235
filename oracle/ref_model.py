"""numpy restatement of the RGCN/PPI whole-model forward (TEST INFRASTRUCTURE; PARITY UNPINNED):
models/sparse_graph_model.py:162-202 with RGCN_Model's defaults + tasks/ppi_task.py:176-179."""
import numpy as np

from . import ref_layers as R


def rgcn_ppi_logits(features, adjacency_lists, type_to_num_incoming_edges, params, weights, dtype=np.float64):
    """weights: {"projection" [F,H] | None, "layers": [ {"edge_weights": L x [H,H]} ... ], "inter_dense": {layer: [H,H]},
    "out_kernel" [H, labels], "out_bias" [labels]}"""
    act = R.get_activation(params["graph_model_activation_function"])
    cur = np.asarray(features, dtype)
    if weights.get("projection") is not None:
        cur = R._apply_act(act, cur @ np.asarray(weights["projection"], dtype))                   # :165-170
    last_residual = np.zeros_like(cur)
    for l in range(params["graph_num_layers"]):
        if l % params["graph_residual_connection_every_num_layers"] == 0:                           # :180-185
            t = cur
            if l > 0:
                cur = (cur + last_residual) / 2
            last_residual = t
        cur = R.sparse_rgcn_layer(cur, adjacency_lists, type_to_num_incoming_edges, params["hidden_size"],
                                  num_timesteps=params["graph_num_timesteps_per_layer"],
                                  activation_function=params["graph_activation_function"],
                                  message_aggregation_function=params["message_aggregation_function"],
                                  weights=weights["layers"][l], dtype=dtype)                       # :187-191
        if l in weights.get("inter_dense", {}):                                                     # :194-200
            cur = R._apply_act(act, cur @ np.asarray(weights["inter_dense"][l], dtype))
    return cur @ np.asarray(weights["out_kernel"], dtype) + np.asarray(weights["out_bias"], dtype)  # ppi_task.py:176-179
