"""CPU: the epoch-loop bookkeeping (training.py) with a stub model -- counters, the reference's log-line formats, the
batch-weighted epoch loss, early stopping with patience and save-best (models/sparse_graph_model.py:263-371)."""
import re

import numpy as np
import pytest

from tf_gnn_samples_b200 import batching, training

torch = pytest.importorskip("torch")


class StubModel:
    """Scaffold interface; 'loss' follows a script so that the early-stopping logic can be exercised."""

    def __init__(self, valid_total_losses):
        self.valid_total_losses = list(valid_total_losses)
        self.train_calls, self.eval_calls, self.mode = 0, 0, "train"
        self._epoch_valid = None

    def train_step_async(self, optimizer, features, plan, num_incoming, targets):
        self.train_calls += 1
        return {"loss": torch.tensor(0.5), "total_loss": torch.tensor(0.5 * features.shape[0]), "f1_score": torch.tensor(0.25)}

    def eval(self):
        self.mode = "eval"

    def __call__(self, features, plan, num_incoming):
        return features

    def task_metrics(self, outputs, targets):
        self.eval_calls += 1
        total = self.valid_total_losses[0]
        return {"loss": torch.tensor(total / outputs.shape[0]), "total_loss": torch.tensor(float(total)), "f1_score": torch.tensor(0.75)}


def ppi_batches(n=3):
    out = []
    for i in range(n):
        b = batching.pack_batch([batching.make_ppi_like_graph(20 + i, 50, feature_dim=4, seed=i), batching.make_ppi_like_graph(10, 30, feature_dim=4, seed=10 + i)])
        out.append(training.TaskBatch(b, np.zeros((b.num_nodes, 3), np.float32)))
    return out


def fake_to_device(tb):
    return (torch.as_tensor(tb.batch.node_features), None, torch.as_tensor(tb.batch.type_to_num_incoming_edges), torch.as_tensor(tb.targets))


def test_run_epoch_counters_and_weighted_loss():
    model = StubModel([12.0])
    ticks = iter([100.0, 102.0])
    loss, metrics, graphs, gps, nps, eps = training.run_epoch(model, None, ppi_batches(), True, fake_to_device, clock=lambda: next(ticks))
    bs = ppi_batches()
    assert graphs == 6 and model.train_calls == 3 and len(metrics) == 3
    assert loss == pytest.approx(0.5)                                        # sum(loss_b * graphs_b) / graphs (:296,309)
    assert gps == pytest.approx(6 / 2.0) and nps == pytest.approx(sum(b.batch.num_nodes for b in bs) / 2.0)
    assert eps == pytest.approx(sum(b.batch.num_edges for b in bs) / 2.0)    # edges = all adjacency rows of all types (:285)
    with pytest.raises(AssertionError):
        training.run_epoch(model, None, [], True, fake_to_device)


def test_metric_summaries_match_the_reference_formats():
    assert training.pretty_print_epoch_task_metrics("PPI", [{"f1_score": 0.5}, {"f1_score": 1.0}], 4) == "Avg MicroF1: 0.750"
    s = training.pretty_print_epoch_task_metrics("QM9", [{"abs_err_task0": 2.0, "abs_err_task4": 1.0}, {"abs_err_task0": 2.0, "abs_err_task4": 3.0}], 8, (0, 4))
    assert s == "MAEs: 0:0.50000 4:0.50000 | Error Ratios: 0:%.5f 4:%.5f" % (0.5 / 0.066513725, 0.5 / 0.033486113)
    assert training.early_stopping_metric([{"total_loss": 6.0}, {"total_loss": 2.0}], 4) == 2.0
    with pytest.raises(ValueError):
        training.pretty_print_epoch_task_metrics("varmisuse", [], 1)


def test_train_loop_logs_saves_best_and_stops_after_patience():
    model = StubModel([9.0])
    script = iter([9.0, 6.0, 7.0, 8.0, 8.5, 1.0])                # validation total loss per epoch (single-batch validation)
    saved, lines = [], []

    def valid():
        model.valid_total_losses = [next(script)]
        return ppi_batches(1)

    res = training.train(model, None, "ppi", lambda: ppi_batches(2), valid, fake_to_device, max_epochs=50, patience=3,
                         log=lines.append, save_best=lambda: saved.append(len(lines)), best_model_file="best.pickle")
    assert res["best_epoch"] == 2 and res["best_valid_metric"] == pytest.approx(6.0 / 2) and len(saved) == 2
    assert len(res["history"]) == 5                               # epochs 3, 4, 5 do not improve -> stop at 5 with patience 3
    assert lines[0] == "== Epoch 1"
    assert re.fullmatch(r" Train: loss: 0\.50000 \|\| Avg MicroF1: 0\.250 \|\| graphs/sec: [\d.]+ \| nodes/sec: \d+ \| edges/sec: \d+", lines[1])
    assert re.fullmatch(r" Valid: loss: [\d.]+ \|\| Avg MicroF1: 0\.750 \|\| graphs/sec: [\d.]+ \| nodes/sec: \d+ \| edges/sec: \d+", lines[2])
    assert lines[3] == "  (Best epoch so far, target metric decreased to 4.50000 from inf. Saving to 'best.pickle')"
    assert lines[-2] == "Stopping training after 3 epochs without improvement on validation loss."
    assert lines[-1].startswith("Training took ") and lines[-1].endswith("Best validation results: Avg MicroF1: 0.750")
    assert model.mode == "eval" and model.eval_calls == 5


def test_prefetch_keeps_order_bounds_the_queue_and_propagates_errors():
    import threading
    import time
    produced, lock = [], threading.Lock()

    def fn(i):
        with lock:
            produced.append(i)
        return i * i

    it = training.prefetch(range(50), fn, max_queue_size=3)
    first = next(it)
    time.sleep(0.2)                                                # the worker runs ahead, but only by the queue size (+1 in flight)
    with lock:
        ahead = len(produced)
    assert first == 0 and 1 <= ahead <= 1 + 3 + 1
    assert [first] + list(it) == [i * i for i in range(50)]

    def bad(i):
        if i == 3:
            raise KeyError("boom")
        return i
    got = []
    with pytest.raises(KeyError):
        for v in training.prefetch(range(10), bad):
            got.append(v)
    assert got == [0, 1, 2]
    assert list(training.prefetch([], fn)) == []


def test_train_step_async_plumbing_on_the_host():
    """SparseGraphModel.train_step_async with the GNN forward replaced by a host computation: learning-rate normalisation,
    backward of task_metrics['loss'], per-tensor clipping and the optimizer step run in the order of __make_train_step."""
    import torch
    from tf_gnn_samples_b200.scaffold import SparseGraphModel

    class HostModel(SparseGraphModel):
        def forward(self, features, plan, num_incoming, graph_nodes_list=None, num_graphs=None):
            final = torch.tanh(features @ self.projection)
            outs = []
            for hd in self.head:
                gate = torch.sigmoid(torch.cat([final, features], dim=-1) @ hd["gate_kernel"] + hd["gate_bias"])
                per_node = gate * (final @ hd["kernel"] + hd["bias"])
                outs.append(torch.zeros(num_graphs, 1).index_add_(0, graph_nodes_list, per_node)[:, 0])
            return torch.stack(outs)

    torch.manual_seed(0)
    params = {"hidden_size": 8, "graph_num_layers": 1, "optimizer": "RMSProp", "learning_rate": 0.01, "lr_for_num_graphs_per_batch": 4,
              "clamp_gradient_norm": 0.05}
    m = HostModel("rgcn", "qm9", 2, 5, params=params, task_ids=(0,), device="cpu")
    opt = m.make_optimizer()
    feats, ids = torch.randn(9, 5), torch.tensor([0, 0, 0, 1, 1, 2, 2, 2, 2])
    targets = torch.randn(1, 3)
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    out = m.train_step_async(opt, feats, None, None, targets, ids, 3)
    assert opt.param_groups[0]["lr"] == pytest.approx(0.01 * 3 / 4)                  # 3 graphs in the batch, normalised to 4
    assert set(out) >= {"loss", "total_loss", "abs_err_task0"} and float(out["total_loss"]) == pytest.approx(3 * float(out["loss"]))
    used = [n for n, p in m.named_parameters() if p.grad is not None]
    assert "projection" in used and all(float(m.get_parameter(n).grad.norm()) <= 0.05 * (1 + 1e-5) for n in used)
    assert any(not torch.equal(p.detach(), before[n]) for n, p in m.named_parameters())
    with pytest.raises(ValueError):
        m.train_step_async(opt, feats, None, None, targets, ids, None)               # the normalisation needs the graph count
