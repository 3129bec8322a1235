"""numpy restatement of the whole-model forward (TEST INFRASTRUCTURE): models/sparse_graph_model.py:162-202, the
models/<x>_model.py adapters, the PPI head (tasks/ppi_task.py:176-195) and the QM9 head (tasks/qm9_task.py:162-199).

PINNED: tests/test_reference_model_pin.py runs the reference's own scaffold + heads (unmodified, eagerly under
tests/tf1_shim.graph_mode) for all seven model classes on both tasks and compares final node representations and task metrics
with this module at 1e-12, the weights taken from the reference's variables by name (fixtures tests/golden/ref_model_*.npz)."""
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


def apply_gnn_layer(kind, cur, adjacency_lists, type_to_num_incoming_edges, params, w, dtype=np.float64):
    """models/<kind>_model.py:_apply_gnn_layer -- the kwargs each adapter forwards to its layer function."""
    H, T = params["hidden_size"], params["graph_num_timesteps_per_layer"]
    act, agg = params["graph_activation_function"], params.get("message_aggregation_function", "sum")
    if kind == "rgcn":                                                                              # rgcn_model.py:36-44
        return R.sparse_rgcn_layer(cur, adjacency_lists, type_to_num_incoming_edges, H, num_timesteps=T, activation_function=act,
                                   message_aggregation_function=agg, weights=w, dtype=dtype)
    if kind == "ggnn":                                                                              # ggnn_model.py:37-45
        return R.sparse_ggnn_layer(cur, adjacency_lists, H, num_timesteps=T, gated_unit_type=params["graph_rnn_cell"],
                                   activation_function=act, message_aggregation_function=agg, weights=w, dtype=dtype)
    if kind == "rgat":                                                                              # rgat_model.py:36-43
        return R.sparse_rgat_layer(cur, adjacency_lists, H, num_timesteps=T, num_heads=params["num_heads"],
                                   activation_function=act, weights=w, dtype=dtype)
    if kind == "gnn-film":                                                                          # gnn_film_model.py:34-43
        return R.sparse_gnn_film_layer(cur, adjacency_lists, type_to_num_incoming_edges, H, num_timesteps=T, activation_function=act,
                                       message_aggregation_function=agg,
                                       normalize_by_num_incoming=params["normalize_messages_by_num_incoming"], weights=w, dtype=dtype)
    if kind == "gnn-edge-mlp":                                                                      # gnn_edge_mlp_model.py:38-48
        return R.sparse_gnn_edge_mlp_layer(cur, adjacency_lists, type_to_num_incoming_edges, H, num_timesteps=T,
                                           activation_function=act, message_aggregation_function=agg,
                                           use_target_state_as_input=params["use_target_state_as_input"],
                                           num_edge_hidden_layers=params["num_edge_hidden_layers"], weights=w, dtype=dtype)
    if kind == "rgin":                                                                              # rgin_model.py:39-49
        return R.sparse_rgin_layer(cur, adjacency_lists, H, num_timesteps=T, activation_function=act,
                                   message_aggregation_function=agg, use_target_state_as_input=params["use_target_state_as_input"],
                                   num_edge_MLP_hidden_layers=params["graph_num_edge_MLP_hidden_layers"],
                                   num_aggr_MLP_hidden_layers=params["graph_num_aggr_MLP_hidden_layers"], weights=w, dtype=dtype)
    if kind == "rgdcn":                                                                             # rgdcn_model.py:34-52
        return R.sparse_rgdcn_layer(cur, adjacency_lists, type_to_num_incoming_edges, params["num_channels"],
                                    H // params["num_channels"], T, params["use_full_state_for_channel_weights"],
                                    params["tie_channel_weights"], act, agg, weights=w, dtype=dtype)
    raise ValueError("Unknown model type '%s'" % kind)


def node_representations(kind, features, adjacency_lists, type_to_num_incoming_edges, params, projection, layers, dtype=np.float64):
    """models/sparse_graph_model.py:162-202 (inference: dropout is the identity).  ``layers``: per layer the layer
    function's weight dict plus optional "inter_ln_gamma"/"inter_ln_beta" and "inter_dense"."""
    act = R.get_activation(params["graph_model_activation_function"])
    cur = np.asarray(features, dtype)
    if projection is not None:
        cur = R._apply_act(act, cur @ np.asarray(projection, dtype))
    last_residual = np.zeros_like(cur)
    for l, w in enumerate(layers):
        if l % params["graph_residual_connection_every_num_layers"] == 0:
            t = cur
            if l > 0:
                cur = (cur + last_residual) / 2
            last_residual = t
        cur = apply_gnn_layer(kind, cur, adjacency_lists, type_to_num_incoming_edges, params, w, dtype)
        if w.get("inter_ln_gamma") is not None:                                                     # :192-193
            cur = R.layer_norm(cur, np.asarray(w["inter_ln_gamma"], dtype), np.asarray(w["inter_ln_beta"], dtype))
        if w.get("inter_dense") is not None:                                                        # :194-200
            cur = R._apply_act(act, cur @ np.asarray(w["inter_dense"], dtype))
    return cur


def qm9_outputs(final, features, graph_nodes_list, num_graphs, heads, dtype=np.float64):
    """tasks/qm9_task.py:162-189: per task sigmoid(gate([h | x0])) * transform(h), summed per graph -> [tasks, G]."""
    final, features = np.asarray(final, dtype), np.asarray(features, dtype)
    gate_in = np.concatenate([final, features], axis=-1)
    outs = []
    for hd in heads:
        per_node = final @ np.asarray(hd["kernel"], dtype) + np.asarray(hd["bias"], dtype)
        gate = 1.0 / (1.0 + np.exp(-(gate_in @ np.asarray(hd["gate_kernel"], dtype) + np.asarray(hd["gate_bias"], dtype))))
        outs.append(R.unsorted_segment_sum(gate * per_node, np.asarray(graph_nodes_list, np.int64), num_graphs)[:, 0])
    return np.stack(outs)


def qm9_metrics(outputs, targets, task_ids):
    """tasks/qm9_task.py:191-199."""
    err = np.asarray(outputs, np.float64) - np.asarray(targets, np.float64)
    m = {"abs_err_task%d" % t: float(np.abs(err[i]).sum()) for i, t in enumerate(task_ids)}
    m["loss"] = float(sum(np.mean(0.5 * err[i] ** 2) for i in range(err.shape[0])))
    m["total_loss"] = m["loss"] * err.shape[1]
    return m


def ppi_metrics(logits, labels):
    """tasks/ppi_task.py:181-195 + utils/utils.py:61-74: summed sigmoid cross-entropy (max(x,0) - x*z + log(1+exp(-|x|))),
    loss = total / number of nodes, micro-F1 from integer counts of round(sigmoid(x)) (round-half-even: 0.5 -> 0)."""
    x, z = np.asarray(logits, np.float64), np.asarray(labels, np.float64)
    total = float(np.sum(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))))
    pred, lab = np.round(1.0 / (1.0 + np.exp(-x))).astype(np.int64), z.astype(np.int64)
    tp, fp, fn = np.count_nonzero(pred * lab), np.count_nonzero(pred * (lab - 1)), np.count_nonzero((pred - 1) * lab)
    precision, recall = tp / (tp + fp), tp / (tp + fn)
    return {"loss": total / x.shape[0], "total_loss": total, "f1_score": float(2 * precision * recall / (precision + recall))}
