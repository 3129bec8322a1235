"""Whole-model cases: the reference's OWN scaffold + task heads (models/sparse_graph_model.py, models/*_model.py,
tasks/ppi_task.py, tasks/qm9_task.py -- unmodified, built eagerly under tests/tf1_shim.graph_mode) against oracle/ref_model.py
fed through tf-gnn-samples_b200/checkpoint.py's variable-name sorting.

run_reference(case, dtype)   -> feed (placeholder name -> array), every variable under the name the reference created it with,
                                final node representations, task metrics, the "Model has N parameters." count
run_oracle(case, feed, vars) -> the same quantities from the numpy oracle, weights taken from the reference's variables BY NAME
Shared by tests/golden/make_model_fixtures.py and tests/test_reference_model_pin.py."""
import importlib
import os
import re
import tempfile

import numpy as np

import batcher_cases as BC

MODEL_CLASSES = {"rgcn": "RGCN_Model", "ggnn": "GGNN_Model", "rgat": "RGAT_Model", "gnn-film": "GNN_FiLM_Model",
                 "gnn-edge-mlp": "GNN_Edge_MLP_Model", "rgin": "RGIN_Model", "rgdcn": "RGDCN_Model"}
_SMALL = {"hidden_size": 32, "graph_num_layers": 3}
_SCAFFOLD = {"graph_residual_connection_every_num_layers": 2, "graph_dense_between_every_num_gnn_layers": 2,
             "graph_inter_layer_norm": True, "graph_model_activation_function": "gelu"}
CASES = {
    # every model class with its own defaults (only made small) on 60-odd real QM9 molecules, two regression tasks
    "rgcn_qm9": dict(kind="rgcn", task="qm9", model_params=_SMALL, task_params={"task_ids": [0, 4]}, budget=600),
    "ggnn_qm9": dict(kind="ggnn", task="qm9", model_params=dict(_SMALL, graph_num_timesteps_per_layer=2), task_params={}, budget=600),
    "rgat_qm9": dict(kind="rgat", task="qm9", model_params=_SMALL, task_params={}, budget=600),
    "film_qm9": dict(kind="gnn-film", task="qm9", model_params=_SMALL, task_params={}, budget=600),
    "edge_mlp_qm9": dict(kind="gnn-edge-mlp", task="qm9", model_params=_SMALL, task_params={}, budget=600),
    "rgin_qm9": dict(kind="rgin", task="qm9", model_params=_SMALL, task_params={}, budget=600),
    "rgdcn_qm9": dict(kind="rgdcn", task="qm9", model_params=_SMALL, task_params={}, budget=600),
    # the scaffold's options switched on: residuals, inter-layer Dense every 2nd layer, inter-layer LayerNorm, two timesteps
    "rgcn_ppi": dict(kind="rgcn", task="ppi", model_params=_SMALL, task_params={}, budget=10 ** 6),
    "rgcn_ppi_scaffold": dict(kind="rgcn", task="ppi", model_params=dict(_SMALL, graph_num_layers=4, graph_num_timesteps_per_layer=2,
                                                                          message_aggregation_function="mean", **_SCAFFOLD),
                              task_params={}, budget=130),
    "film_ppi_scaffold": dict(kind="gnn-film", task="ppi", model_params=dict(_SMALL, graph_num_layers=4, graph_num_timesteps_per_layer=2,
                                                                             normalize_messages_by_num_incoming=True, **_SCAFFOLD),
                              task_params={"tie_fwd_bkwd_edges": True}, budget=10 ** 6),
    "ggnn_ppi_hidden_is_feature_size": dict(kind="ggnn", task="ppi", model_params={"hidden_size": 7, "graph_num_layers": 2},
                                            task_params={}, budget=10 ** 6),     # no input projection: the head is 'dense', not 'dense_1'
    "rgin_ppi_scaffold": dict(kind="rgin", task="ppi", model_params=dict(_SMALL, graph_num_timesteps_per_layer=2,
                                                                         graph_num_aggr_MLP_hidden_layers=1, **_SCAFFOLD),
                              task_params={}, budget=10 ** 6),
}
README_RGCN_PPI = dict(kind="rgcn", model_params={"hidden_size": 256, "graph_num_layers": 3}, feature_dim=50, num_labels=121)


def _first_feed(tf1_shim, case, data_dir):
    """Phase 1 (inert placeholders): the task object, loaded through the reference's loader, and its first minibatch."""
    from dpu_utils.utils import RichPath
    sgt = tf1_shim.import_reference_task("sparse_graph_task")
    if case["task"] == "qm9":
        mod = tf1_shim.import_reference_task("qm9_task")
        cls, fold, target, path = mod.QM9_Task, sgt.DataFold.VALIDATION, "target_values", os.path.join(data_dir, "valid.jsonl.gz")
    else:
        mod = tf1_shim.import_reference_task("ppi_task")
        cls, fold, target, path = mod.PPI_Task, sgt.DataFold.TEST, "target_labels", data_dir
    task_params = cls.default_params()
    task_params.update(case["task_params"])
    task = cls(task_params)
    data = task.load_eval_data_from_path(RichPath.create(path))
    names = ["initial_node_features", "type_to_num_incoming_edges", "graph_nodes_list", target, "out_layer_dropout_keep_prob"]
    feed = BC._feeds_of(task, data, fold, names, case["budget"])[0]
    return task, feed


def prepare_data_dir(case, data_dir, ppi_kw=None):
    import shutil
    if case["task"] == "qm9":
        shutil.copy(BC.QM9_SUBSET, os.path.join(data_dir, "valid.jsonl.gz"))
    else:
        BC.write_ppi_dir(data_dir, "test", **(ppi_kw or {}))


def run_reference(case, dtype=np.float64, seed=11, ppi_kw=None, provider=None, gradient_hook=None):
    """``provider``: explicit variable values by tf name (tf1_shim.variables.provider_from); default: seeded initialisers."""
    import tf1_shim
    with tempfile.TemporaryDirectory() as tmp, tf1_shim.installed(dtype=dtype, seed=seed, provider=provider) as session:
        prepare_data_dir(case, tmp, ppi_kw)
        task, feed = _first_feed(tf1_shim, case, tmp)
        session.feeds = feed                                        # phase 2: placeholders hand out the feed, the graph runs as it is built
        session.gradient_hook = gradient_hook
        import models                                               # the reference's package
        cls = getattr(models, MODEL_CLASSES[case["kind"]])
        params = cls.default_params()
        params.update(case["model_params"])
        model = cls(params, task, "run", tmp)
        ops = model._Sparse_Graph_Model__ops
        with open(model.log_file) as f:
            count = int(re.search(r"Model has (\d+) parameters", f.read()).group(1))
        snapshot = os.path.join(tmp, "snapshot.pickle")
        model.save_model(snapshot)                                  # the reference's own pickle (sparse_graph_model.py:91-107)
        with open(snapshot, "rb") as f:
            pickled = f.read()
        return {"feed": feed, "variables": dict(session.variables), "params": dict(model.params), "task_params": dict(task.params),
                "final": np.asarray(ops["final_node_representations"]),
                "metrics": {k: np.asarray(v) for k, v in ops["task_metrics"].items()},
                "num_parameters": count, "num_edge_types": task.num_edge_types, "pickle": pickled,
                "model_name": cls.name(params), "task_name": task.name(),
                "optimizers": list(session.optimizers), "applied": session.applied,
                "loss_is_task_loss": session.loss_for_gradients is ops["task_metrics"]["loss"]}


def adjacency_of(feed, num_edge_types):
    return [np.asarray(feed["adjacency_e%d" % i]).reshape(-1, 2) for i in range(num_edge_types)]


def run_oracle(case, feed, variables, params, task_params, num_edge_types, dtype=np.float64):
    from oracle import ref_model
    ck = importlib.import_module("tf_gnn_samples_b200.checkpoint")
    scaffold = importlib.import_module("tf_gnn_samples_b200.scaffold")
    srt = ck.sort_variables(variables)
    assert not srt["unused"], srt["unused"]
    kind = case["kind"]
    T = params["graph_num_timesteps_per_layer"] if kind in scaffold._LAYERS_WITH_OWN_LN else 0
    layers = [ck.split_layer_norms(l, T) for l in srt["layers"]]
    assert srt["layer_indices"] == list(range(params["graph_num_layers"]))
    feats = np.asarray(feed["initial_node_features"]).astype(np.float32).astype(dtype)      # fp32 placeholder (sparse_graph_task.py:139)
    outside = ck.scaffold_variables(srt["outside"], feats.shape[1], params["hidden_size"])
    adj = adjacency_of(feed, num_edge_types)
    indeg = np.asarray(feed["type_to_num_incoming_edges"]).astype(np.float32).astype(dtype)
    final = ref_model.node_representations(kind, feats, adj, indeg, params, outside.get("projection"), layers, dtype=dtype)
    if case["task"] == "ppi":
        head = [h for h in outside["head"] if "bias" in h][-1]
        logits = final @ np.asarray(head["kernel"], dtype) + np.asarray(head["bias"], dtype)
        metrics = ref_model.ppi_metrics(logits, feed["target_labels"])
    else:
        task_ids = task_params["task_ids"]
        heads = [outside["qm9_heads"][t] for t in task_ids]
        outs = ref_model.qm9_outputs(final, feats, feed["graph_nodes_list"], int(feed["num_graphs"]), heads, dtype=dtype)
        metrics = ref_model.qm9_metrics(outs, np.asarray(feed["target_values"]).astype(np.float32), task_ids)   # fp32 placeholder (qm9_task.py:157)
    return {"final": final, "metrics": metrics, "layers": layers, "outside": outside}
