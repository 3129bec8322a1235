"""Shared by tests/test_reference_batcher_pin.py and tests/golden/make_batcher_fixtures.py: run the REFERENCE's own task
loaders + minibatch iterators (tasks/qm9_task.py, tasks/ppi_task.py, unmodified, under tests/tf1_shim) and this repo's
batching.py on the same files, and bring both feeds into one comparable form.

A feed (one minibatch) is a dict  name -> array  with the reference's placeholder names: initial_node_features,
type_to_num_incoming_edges, graph_nodes_list, target_values | target_labels, adjacency_e<i>, plus num_graphs / num_nodes /
num_edges of the MinibatchData tuple (tasks/sparse_graph_task.py:15-19)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
QM9_SUBSET = os.path.join(HERE, "qm9_valid_subset.json.gz")

QM9_CASES = {
    # name: (task params on top of QM9_Task.default_params(), max_nodes_per_batch)
    "qm9_default": ({}, 400),
    "qm9_two_tasks_budget_90": ({"task_ids": [3, 7]}, 90),
    "qm9_no_self_loops": ({"add_self_loop_edges": False}, 5000),
}
# tie_fwd_bkwd_edges=False cannot run in the reference's QM9 loader: qm9_task.py:139-145 appends the backward lists to the
# very list it is enumerating, so the loop runs past num_edge_types and raises IndexError on the first molecule.
QM9_REFERENCE_RAISES = {
    "qm9_untied": ({"tie_fwd_bkwd_edges": False}, 1000),
    "qm9_untied_no_self_loops": ({"add_self_loop_edges": False, "tie_fwd_bkwd_edges": False}, 5000),
}
PPI_CASES = {
    "ppi_default": ({}, 130),
    "ppi_tied": ({"tie_fwd_bkwd_edges": True}, 1000),
    "ppi_no_self_loops": ({"add_self_loop_edges": False}, 70),
    "ppi_tied_no_self_loops": ({"add_self_loop_edges": False, "tie_fwd_bkwd_edges": True}, 10 ** 6),
}


def write_ppi_dir(path, fold="valid", seed=0, num_graphs=5, feature_dim=7, num_labels=4, linkless_graph=None):
    """A PPI fold in the layout of the public dgl download the reference reads (tasks/ppi_task.py:69,85-88): <fold>_graph.json
    with a node-link 'links' list over GLOBAL node ids, <fold>_feats.npy, <fold>_labels.npy, <fold>_graph_id.npy.  Graph ids
    do not start at 0 and are not consecutive (as in the real data: 1..20 train, 21.. valid).  ``linkless_graph``: index of a
    graph without any link -- the reference then builds np.array([]) of shape (0,) for it (ppi_task.py:152) and
    np.concatenate with the (E, 2) lists of its batch neighbours raises ValueError (:247); batching.py keeps (0, 2)."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(12, 60, size=num_graphs)
    ids = 21 + 2 * np.arange(num_graphs)
    graph_id = np.repeat(ids, sizes)
    n = int(sizes.sum())
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    links = []
    for g in range(num_graphs):
        if g == linkless_graph:
            continue
        m = int(rng.integers(20, 150))
        s = starts[g] + rng.integers(0, sizes[g], size=m)
        t = starts[g] + rng.integers(0, sizes[g], size=m)
        links += [{"source": int(a), "target": int(b)} for a, b in zip(s, t)]
    order = rng.permutation(len(links))          # file order is not grouped by graph
    links = [links[i] for i in order]
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "%s_graph.json" % fold), "w") as f:
        json.dump({"directed": False, "multigraph": False, "nodes": [{"id": i} for i in range(n)], "links": links}, f)
    np.save(os.path.join(path, "%s_feats.npy" % fold), rng.standard_normal((n, feature_dim)))
    np.save(os.path.join(path, "%s_labels.npy" % fold), (rng.random((n, num_labels)) < 0.3).astype(np.int64))
    np.save(os.path.join(path, "%s_graph_id.npy" % fold), graph_id)
    return path


def _feeds_of(task, data, fold, names, max_nodes):
    import tensorflow as tf        # the shim (inside tf1_shim.installed())
    ph = {k: tf.placeholder(None, name=k) for k in names}
    ph["adjacency_lists"] = [tf.placeholder(None, name="adjacency_e%d" % i) for i in range(task.num_edge_types)]
    feeds = []
    for mb in task.make_minibatch_iterator(data, fold, ph, max_nodes):
        feed = {k.name: np.asarray(v) for k, v in mb.feed_dict.items()}
        feed.update(num_graphs=np.int64(mb.num_graphs), num_nodes=np.int64(mb.num_nodes), num_edges=np.int64(mb.num_edges))
        feeds.append(feed)
    return feeds


def reference_qm9_feeds(params, max_nodes, path=QM9_SUBSET):
    """QM9_Task.load_eval_data_from_path + make_minibatch_iterator (validation fold: no shuffling) on the 200 real molecules
    of tests/golden/qm9_valid_subset.json.gz, copied to a *.jsonl.gz name as RichPath dispatches on the suffix."""
    import shutil
    import tempfile
    import tf1_shim
    with tempfile.TemporaryDirectory() as tmp, tf1_shim.installed():
        from dpu_utils.utils import RichPath
        mod = tf1_shim.import_reference_task("qm9_task")
        sgt = tf1_shim.import_reference_task("sparse_graph_task")
        task_params = mod.QM9_Task.default_params()
        task_params.update(params)
        task = mod.QM9_Task(task_params)
        file = os.path.join(tmp, "valid.jsonl.gz")
        shutil.copy(path, file)
        data = task.load_eval_data_from_path(RichPath.create(file))
        names = ["initial_node_features", "type_to_num_incoming_edges", "graph_nodes_list", "target_values",
                 "out_layer_dropout_keep_prob"]
        return _feeds_of(task, data, sgt.DataFold.VALIDATION, names, max_nodes), task.num_edge_types


def reference_ppi_feeds(params, max_nodes, data_dir):
    """PPI_Task.load_eval_data_from_path (the 'test' fold) + make_minibatch_iterator."""
    import tf1_shim
    with tf1_shim.installed():
        from dpu_utils.utils import RichPath
        mod = tf1_shim.import_reference_task("ppi_task")
        sgt = tf1_shim.import_reference_task("sparse_graph_task")
        task_params = mod.PPI_Task.default_params()
        task_params.update(params)
        task = mod.PPI_Task(task_params)
        data = task.load_eval_data_from_path(RichPath.create(data_dir))
        names = ["initial_node_features", "type_to_num_incoming_edges", "graph_nodes_list", "target_labels",
                 "out_layer_dropout_keep_prob"]
        return _feeds_of(task, data, sgt.DataFold.TEST, names, max_nodes), task.num_edge_types


def repo_feed(batch, extra):
    """A batching.Batch (+ task arrays) under the reference's placeholder names."""
    sizes = np.diff(batch.graph_node_offsets)
    feed = {"initial_node_features": batch.node_features,
            "type_to_num_incoming_edges": batch.type_to_num_incoming_edges,
            "graph_nodes_list": np.repeat(np.arange(batch.num_graphs, dtype=np.int32), sizes),
            "num_graphs": np.int64(batch.num_graphs), "num_nodes": np.int64(batch.num_nodes),
            "num_edges": np.int64(batch.num_edges)}
    for i, a in enumerate(batch.adjacency_lists):
        feed["adjacency_e%d" % i] = a
    feed.update(extra)
    return feed


def repo_qm9_feeds(params, max_nodes, path=QM9_SUBSET):
    import importlib
    batching = importlib.import_module("tf_gnn_samples_b200.batching")
    self_loops, tied = params.get("add_self_loop_edges", True), params.get("tie_fwd_bkwd_edges", True)
    task_ids = params.get("task_ids", [0])
    recs = batching.load_qm9_jsonl(path)
    L = batching.qm9_num_edge_types(recs, self_loops, tied)
    samples = [batching.qm9_graph_to_sample(r, L, self_loops, tied) for r in recs]
    feeds = []
    for batch, first in batching.minibatches(samples, max_nodes):
        graphs = recs[first:first + batch.num_graphs]
        targets = np.array([[g["targets"][t][0] for g in graphs] for t in task_ids])
        feeds.append(repo_feed(batch, {"target_values": targets}))
    return feeds, L


def repo_ppi_feeds(params, max_nodes, data_dir):
    import importlib
    batching = importlib.import_module("tf_gnn_samples_b200.batching")
    graphs, labels = batching.load_ppi_fold(data_dir, "test", params.get("add_self_loop_edges", True),
                                            params.get("tie_fwd_bkwd_edges", False))
    feeds = []
    for batch, first in batching.minibatches(graphs, max_nodes):
        feeds.append(repo_feed(batch, {"target_labels": np.concatenate(labels[first:first + batch.num_graphs], axis=0)}))
    return feeds, len(graphs[0].adjacency_lists)


def compare_feeds(got, want, what):
    """Integer arrays (adjacency, graph ids, counts) bit-exact incl. ORDER; float arrays equal after the float32 cast of the
    placeholder (tasks/sparse_graph_task.py:139-146 feed fp32; the reference's numpy arrays are float64 / int)."""
    assert len(got) == len(want), "%s: %d minibatches, reference made %d" % (what, len(got), len(want))
    for i, (g, w) in enumerate(zip(got, want)):
        w = {k: v for k, v in w.items() if k != "out_layer_dropout_keep_prob"}
        assert set(g) == set(w), "%s batch %d: %s" % (what, i, sorted(set(g) ^ set(w)))
        for k in w:
            a, b = np.asarray(g[k]), np.asarray(w[k])
            if k.startswith("adjacency_e"):
                a, b = a.reshape(-1, 2), b.reshape(-1, 2)       # the reference yields shape (0,) for an empty PPI list
            assert a.shape == b.shape, "%s batch %d %s: shape %s vs %s" % (what, i, k, a.shape, b.shape)
            if k.startswith("adjacency_e") or k in ("graph_nodes_list", "num_graphs", "num_nodes", "num_edges"):
                assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), "%s batch %d %s" % (what, i, k)
            else:
                assert np.array_equal(a.astype(np.float32), b.astype(np.float32)), "%s batch %d %s" % (what, i, k)


def pack_feeds(feeds):
    """list of feeds -> flat dict for np.savez (keys b<i>/<name>); unpack_feeds is the inverse."""
    out = {"num_batches": np.int64(len(feeds))}
    for i, f in enumerate(feeds):
        for k, v in f.items():
            if k == "out_layer_dropout_keep_prob":
                continue
            v = np.asarray(v)
            if v.dtype == np.float64:
                v = v.astype(np.float32)            # what the fp32 placeholder receives
            out["b%d/%s" % (i, k)] = v
    return out


def unpack_feeds(z, case):
    """Feeds of one case out of ref_batcher_feeds.npz (keys <case>/b<i>/<name>)."""
    feeds = [dict() for _ in range(int(z[case + "/num_batches"]))]
    for k in z.files:
        parts = k.split("/")
        if parts[0] == case and len(parts) == 3:
            feeds[int(parts[1][1:])][parts[2]] = z[k]
    return feeds
