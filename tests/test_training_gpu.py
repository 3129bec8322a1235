"""GPU: the epoch loop of training.py (Sparse_Graph_Model.train / __run_epoch, models/sparse_graph_model.py:263-371) run
against the REAL scaffold models on the engine -- round 1 only exercised it with a stub.

* RGCN / PPI head on synthetic PPI-shaped graphs whose labels are a function of the node features: a few epochs of
  train + validation through the prefetch thread, the reference's log lines, save-best, decreasing validation loss;
* GGNN / QM9 head on the 200 real validation molecules (tests/golden/qm9_valid_subset.json.gz), minibatched like
  tasks/qm9_task.py:200-261.
"""
import os
import re

import numpy as np
import pytest

from tf_gnn_samples_b200 import batching, training
from tf_gnn_samples_b200.scaffold import RGCNPPIModel, SparseGraphModel

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def ppi_task_batches(num_batches, seed, label_map):
    out = []
    for i in range(num_batches):
        b = batching.pack_batch([batching.make_ppi_like_graph(180 + 10 * j, 2500, feature_dim=50, seed=seed + 10 * i + j) for j in range(2)])
        labels = (b.node_features @ label_map > 0).astype(np.float32)           # learnable from the input projection alone
        out.append(training.TaskBatch(b, labels))
    return out


def test_epoch_loop_trains_the_rgcn_ppi_model(cuda_device):
    import torch
    torch.manual_seed(0)
    label_map = np.random.default_rng(5).standard_normal((50, 121)).astype(np.float32)
    train_set, valid_set = ppi_task_batches(4, 100, label_map), ppi_task_batches(2, 900, label_map)
    model = RGCNPPIModel(device=cuda_device, params={"hidden_size": 64, "learning_rate": 0.005})
    opt = model.make_optimizer()
    lines, saved = [], []
    to_dev = lambda tb: training.device_args(tb, cuda_device)   # noqa: E731
    res = training.train(model, opt, "ppi",
                         train_batches=lambda: training.prefetch(train_set, max_queue_size=2),
                         valid_batches=lambda: valid_set, to_device=to_dev, max_epochs=6, patience=10, log=lines.append,
                         save_best=lambda: saved.append({k: v.detach().clone() for k, v in model.state_dict().items()}))
    hist = res["history"]
    assert len(hist) == 6 and all(np.isfinite(h["train_loss"]) and np.isfinite(h["valid_loss"]) for h in hist)
    assert hist[-1]["valid_loss"] < hist[0]["valid_loss"], [h["valid_loss"] for h in hist]           # it learns
    assert hist[-1]["train_loss"] < hist[0]["train_loss"]
    assert res["best_epoch"] >= 1 and len(saved) >= 1 and res["best_valid_metric"] == min(h["valid_metric"] for h in hist)
    assert all(h["train_edges_per_sec"] > 0 and h["valid_edges_per_sec"] > 0 for h in hist)
    # the reference's log-line formats (models/sparse_graph_model.py:340,356 / tasks/ppi_task.py:262-264)
    assert lines[0] == "== Epoch 1"
    pat = re.compile(r"^ (Train|Valid): loss: \d+\.\d{5} \|\| Avg MicroF1: (\d\.\d{3}|nan) \|\| graphs/sec: \d+\.\d{2} \| nodes/sec: \d+ \| edges/sec: \d+$")
    assert sum(1 for ln in lines if pat.match(ln)) == 12, lines[:4]
    assert any(ln.startswith("  (Best epoch so far, target metric decreased to") for ln in lines)


def test_epoch_loop_trains_ggnn_on_real_qm9_molecules(cuda_device):
    import torch
    torch.manual_seed(1)
    recs = batching.load_qm9_jsonl(os.path.join(HERE, "golden", "qm9_valid_subset.json.gz"))
    assert len(recs) == 200

    def batches(rs, size):
        out = []
        for i in range(0, len(rs), size):
            b, gl, tg = batching.qm9_batch(rs[i:i + size], task_ids=(0,))
            out.append(training.TaskBatch(b, tg, gl))
        return out

    train_set, valid_set = batches(recs[:160], 40), batches(recs[160:], 40)
    params = {"graph_num_layers": 2, "hidden_size": 64, "graph_num_timesteps_per_layer": 2, "graph_rnn_cell": "GRU",
              "graph_layer_input_dropout_keep_prob": 1.0, "learning_rate": 0.003}
    model = SparseGraphModel("ggnn", "qm9", num_edge_types=5, feature_size=15, params=params, task_ids=(0,), device=cuda_device)
    opt = model.make_optimizer()
    lines = []
    res = training.train(model, opt, "qm9", train_batches=lambda: training.prefetch(train_set), valid_batches=lambda: valid_set,
                         to_device=lambda tb: training.device_args(tb, cuda_device), max_epochs=5, patience=10, log=lines.append,
                         task_ids=(0,))
    hist = res["history"]
    assert len(hist) == 5 and all(np.isfinite(h["valid_loss"]) for h in hist)
    assert hist[-1]["train_loss"] < hist[0]["train_loss"], [h["train_loss"] for h in hist]
    assert any(re.match(r"^ Valid: loss: \d+\.\d{5} \|\| MAEs: 0:\d+\.\d{5} \| Error Ratios: 0:\d+\.\d{5} \|\| graphs/sec", ln) for ln in lines), lines[:5]
