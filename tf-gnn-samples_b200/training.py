"""Host-side epoch loop around the scaffold models: the bookkeeping of Sparse_Graph_Model.__run_epoch / train
(models/sparse_graph_model.py:263-371) and the tasks' metric summaries (tasks/ppi_task.py:258-264,
tasks/qm9_task.py:263-282) -- same counters (graphs / nodes / edges per second, edges = sum over edge types of the batch's
adjacency rows: :285,310), same log lines (:340,356,361-370), same early stopping on total_loss / num_graphs with save-best
and patience.  Data loading, the CLI and TensorBoard summaries of the reference stay out of scope.

A model is anything with the scaffold's interface (scaffold.SparseGraphModel / RGCNPPIModel): ``train_step_async(optimizer,
features, plan, num_incoming, targets, ...)``, ``__call__``, ``task_metrics``, ``eval()``.  Batches are ``TaskBatch`` records
built on the host (batching.py); ``to_device`` turns one into the argument tuple of the model (GraphPlan, device tensors).
"""
import time
from typing import Any, Callable, Dict, Iterable, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from .batching import Batch

# tasks/qm9_task.py:22-26
QM9_CHEMICAL_ACC_NORMALISING_FACTORS = [0.066513725, 0.012235489, 0.071939046, 0.033730778, 0.033486113, 0.004278493,
                                        0.001330901, 0.004165489, 0.004128926, 0.00409976, 0.004527465, 0.012292586,
                                        0.037467458]


class TaskBatch(NamedTuple):
    batch: Batch                                   # node features, adjacency lists, in-degrees, counters
    targets: np.ndarray                            # PPI: labels [V, num_labels]; QM9: target values [tasks, G]
    graph_nodes_list: Optional[np.ndarray] = None  # QM9: graph id of every node


def device_args(tb: TaskBatch, device) -> Tuple:
    """(features, plan, num_incoming, targets[, graph_nodes_list, num_graphs]) on ``device`` -- what feed_dict carries
    (tasks/sparse_graph_task.py:139-149, tasks/qm9_task.py:242-249)."""
    import torch
    from .engine import GraphPlan
    b = tb.batch
    args = (torch.as_tensor(b.node_features).to(device), GraphPlan(b.adjacency_lists, b.num_nodes, device=device),
            torch.as_tensor(b.type_to_num_incoming_edges).to(device), torch.as_tensor(tb.targets).to(device))
    if tb.graph_nodes_list is not None:
        args += (torch.as_tensor(tb.graph_nodes_list).to(device), b.num_graphs)
    return args


def prefetch(items: Iterable, fn: Callable[[Any], Any] = lambda x: x, max_queue_size: int = 5) -> Iterable:
    """dpu_utils ThreadedIterator as the reference uses it (models/sparse_graph_model.py:270-272, queue of 5): a background
    thread applies ``fn`` (batch construction, host->device copies, GraphPlan build) to the items in order while the
    consumer works on earlier ones.  Exceptions raised by ``fn`` or the source iterator surface in the consumer."""
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=max(1, int(max_queue_size)))
    done = object()

    def worker():
        try:
            for it in items:
                q.put((True, fn(it)))
            q.put((True, done))
        except BaseException as exc:                   # hand the failure to the consumer instead of dying silently
            q.put((False, exc))

    t = threading.Thread(target=worker, daemon=True)
    t.start()
    while True:
        ok, value = q.get()
        if not ok:
            raise value
        if value is done:
            break
        yield value
    t.join()


def pretty_print_epoch_task_metrics(task: str, task_metric_results: List[Dict[str, float]], num_graphs: int,
                                    task_ids: Sequence[int] = (0,)) -> str:
    if task.lower() == "ppi":                                                   # tasks/ppi_task.py:262-264
        # micro_f1 casts to float32 (utils/utils.py:74): the reference averages float32 scalars IN float32
        return "Avg MicroF1: %.3f" % (np.average(np.asarray([m["f1_score"] for m in task_metric_results], dtype=np.float32)),)
    if task.lower() == "qm9":                                                   # tasks/qm9_task.py:267-282
        maes = {t: sum(m["abs_err_task%i" % t] for m in task_metric_results) / float(num_graphs) for t in task_ids}
        maes_str = " ".join("%i:%.5f" % (t, maes[t]) for t in task_ids)
        err_str = " ".join("%i:%.5f" % (t, maes[t] / QM9_CHEMICAL_ACC_NORMALISING_FACTORS[t]) for t in task_ids)
        return "MAEs: %s | Error Ratios: %s" % (maes_str, err_str)
    raise ValueError("Unknown task type '%s'" % task)


def early_stopping_metric(task_metric_results: List[Dict[str, float]], num_graphs: int) -> float:
    """Both tasks stop on the average total loss (tasks/ppi_task.py:258-260, tasks/qm9_task.py:263-265)."""
    return float(np.sum([m["total_loss"] for m in task_metric_results]) / num_graphs)


def run_epoch(model, optimizer, batches: Iterable[TaskBatch], is_training: bool, to_device: Callable[[TaskBatch], Tuple],
              epoch_name: str = "epoch", quiet: bool = True, clock: Callable[[], float] = time.time):
    """models/sparse_graph_model.py:263-316.  Returns (per_graph_loss, task_metric_results, processed_graphs,
    graphs_per_sec, nodes_per_sec, edges_per_sec).  The epoch loss weights every batch's mean loss by its number of graphs
    (:296), exactly like the reference."""
    import torch
    task_metric_results: List[Dict[str, float]] = []
    start = clock()
    graphs = nodes = edges = 0
    epoch_loss = 0.0
    for step, tb in enumerate(batches):
        graphs += tb.batch.num_graphs
        nodes += tb.batch.num_nodes
        edges += tb.batch.num_edges
        args = to_device(tb)
        if is_training:
            m = model.train_step_async(optimizer, *args)
        else:
            model.eval()
            with torch.no_grad():
                m = model.task_metrics(model(*args[:3], *args[4:]), args[3])
        m = {k: float(v) for k, v in m.items()}                # one host read per batch (the reference's sess.run fetch)
        epoch_loss += m["loss"] * tb.batch.num_graphs
        task_metric_results.append(m)
        if not quiet:
            print("Running %s, batch %i (has %i graphs). Loss so far: %.4f" % (epoch_name, step, tb.batch.num_graphs, epoch_loss / graphs), end="\r")
    assert graphs > 0, "Can't run epoch over empty dataset."
    dt = max(clock() - start, 1e-12)
    return epoch_loss / graphs, task_metric_results, graphs, graphs / dt, nodes / dt, edges / dt


def train(model, optimizer, task: str, train_batches: Callable[[], Iterable[TaskBatch]],
          valid_batches: Callable[[], Iterable[TaskBatch]], to_device: Callable[[TaskBatch], Tuple], max_epochs: int = 10000,
          patience: int = 25, log: Callable[[str], None] = print, save_best: Optional[Callable[[], None]] = None,
          best_model_file: str = "<memory>", task_ids: Sequence[int] = (0,), clock: Callable[[], float] = time.time) -> Dict[str, Any]:
    """models/sparse_graph_model.py:323-371: epochs of train + validation, log lines in the reference's format, save on
    improvement of the early-stopping metric, stop after ``patience`` epochs without improvement."""
    total_start = clock()
    best_metric, best_epoch, best_descr = float("+inf"), 0, ""
    history = []
    for epoch in range(1, max_epochs + 1):
        log("== Epoch %i" % epoch)
        tl, tm, tg, tgs, tns, tes = run_epoch(model, optimizer, train_batches(), True, to_device, "epoch %i (training)" % epoch, clock=clock)
        log(" Train: loss: %.5f || %s || graphs/sec: %.2f | nodes/sec: %.0f | edges/sec: %.0f"
            % (tl, pretty_print_epoch_task_metrics(task, tm, tg, task_ids), tgs, tns, tes))
        vl, vm, vg, vgs, vns, ves = run_epoch(model, optimizer, valid_batches(), False, to_device, "epoch %i (validation)" % epoch, clock=clock)
        stop_metric = early_stopping_metric(vm, vg)
        descr = pretty_print_epoch_task_metrics(task, vm, vg, task_ids)
        log(" Valid: loss: %.5f || %s || graphs/sec: %.2f | nodes/sec: %.0f | edges/sec: %.0f" % (vl, descr, vgs, vns, ves))
        history.append({"epoch": epoch, "train_loss": tl, "valid_loss": vl, "valid_metric": stop_metric,
                        "train_edges_per_sec": tes, "valid_edges_per_sec": ves})
        if stop_metric < best_metric:
            if save_best is not None:
                save_best()
            log("  (Best epoch so far, target metric decreased to %.5f from %.5f. Saving to '%s')" % (stop_metric, best_metric, best_model_file))
            best_metric, best_epoch, best_descr = stop_metric, epoch, descr
        elif epoch - best_epoch >= patience:
            log("Stopping training after %i epochs without improvement on validation loss." % patience)
            log("Training took %is. Best validation results: %s" % (clock() - total_start, best_descr))
            break
    return {"best_valid_metric": best_metric, "best_epoch": best_epoch, "best_description": best_descr, "history": history}


def test(model, task: str, batches: Iterable[TaskBatch], to_device: Callable[[TaskBatch], Tuple], data_description: str = "",
         log: Callable[[str], None] = print, task_ids: Sequence[int] = (0,), clock: Callable[[], float] = time.time) -> Dict[str, Any]:
    """Sparse_Graph_Model.test (models/sparse_graph_model.py:373-385): one evaluation epoch over ``batches`` and the three log
    lines of the reference.  (test.py:27 doubles max_nodes_in_batch for it -- the caller's choice when building ``batches``.)"""
    log("== Running Test on %s ==" % (data_description,))
    loss, metrics, num_graphs, _, _, _ = run_epoch(model, None, batches, False, to_device, "Test", clock=clock)
    log("Loss %.5f on %i graphs" % (loss, num_graphs))
    descr = pretty_print_epoch_task_metrics(task, metrics, num_graphs, task_ids)
    log("Metrics: %s" % descr)
    return {"loss": loss, "num_graphs": num_graphs, "description": descr, "task_metric_results": metrics}

