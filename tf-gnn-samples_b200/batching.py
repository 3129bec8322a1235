"""Batch construction: the host-side tensor contract the hot path consumes.

Mirrors the reference's task batchers (tasks/ppi_task.py:197-256, tasks/qm9_task.py:200-261,
tasks/varmisuse_task.py:451-538): graphs are packed into one block-diagonal graph by offsetting
node ids, per-type adjacency lists are concatenated, in-degrees are concatenated along axis 1.
The reference's datasets are not shipped (data/ppi, data/varmisuse) or not available on the GPU
box (data/qm9), so the generators below produce seeded, shape-matched synthetic graphs
(SURVEY.md 8d / Appendix B).  Everything here is numpy on the host, like the reference.
"""
from typing import Dict, Iterator, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np


class GraphSample(NamedTuple):
    """tasks/ppi_task.py:19-23 (without the labels, which belong to the task head, not the hot path)."""
    adjacency_lists: List[np.ndarray]                      # L x int32 [E_l, 2], (src, tgt), node ids local to the graph
    type_to_node_to_num_incoming_edges: np.ndarray         # [L, V_graph]
    node_features: np.ndarray                              # [V_graph, D0] float32


class Batch(NamedTuple):
    """What a reference MinibatchData feed_dict carries for the GNN layers (tasks/sparse_graph_task.py:139-149)."""
    node_features: np.ndarray                              # float32 [V, D0]
    adjacency_lists: List[np.ndarray]                      # L x int32 [E_l, 2]
    type_to_num_incoming_edges: np.ndarray                 # float32 [L, V]
    num_graphs: int
    num_nodes: int
    num_edges: int                                         # sum_l E_l: the reference's edges/sec counter (sparse_graph_model.py:285)
    graph_node_offsets: np.ndarray                         # int64 [num_graphs + 1]


def _in_degrees(adj: Sequence[np.ndarray], num_nodes: int) -> np.ndarray:
    return np.stack([np.bincount(a[:, 1], minlength=num_nodes) if a.shape[0] else np.zeros(num_nodes, np.int64)
                     for a in adj]).astype(np.int32)


def make_ppi_like_graph(num_nodes: int = 2245, num_links: int = 59000, feature_dim: int = 50, seed: int = 0,
                        zipf_targets: bool = False) -> GraphSample:
    """One PPI-shaped graph with the reference's three edge types: 0 = fwd (u,v), 1 = self-loop (i,i),
    2 = bkwd (v,u)  (tasks/ppi_task.py:99-106,125-127,144-148; add_self_loop_edges=True,
    tie_fwd_bkwd_edges=False).  Links are i.i.d. uniform (or Zipf(1.0)-skewed targets)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, num_nodes, size=num_links, dtype=np.int64)
    if zipf_targets:
        p = 1.0 / np.arange(1, num_nodes + 1)
        p /= p.sum()
        tgt = rng.choice(num_nodes, size=num_links, p=p).astype(np.int64)
        tgt = rng.permutation(num_nodes)[tgt]          # hubs are not the low ids
    else:
        tgt = rng.integers(0, num_nodes, size=num_links, dtype=np.int64)
    fwd = np.stack([src, tgt], axis=1).astype(np.int32)
    loops = np.stack([np.arange(num_nodes), np.arange(num_nodes)], axis=1).astype(np.int32)
    bkwd = np.stack([tgt, src], axis=1).astype(np.int32)
    adj = [fwd, loops, bkwd]
    feats = rng.standard_normal((num_nodes, feature_dim)).astype(np.float32)
    return GraphSample(adj, _in_degrees(adj, num_nodes), feats)


def make_qm9_like_graph(rng: np.random.Generator, feature_dim: int = 15, add_self_loop_edges: bool = False) -> GraphSample:
    """One molecule-shaped graph (QM9 statistics measured in SURVEY.md Appendix B: 3..29 nodes, mean 18,
    ~1.03 bonds per node, bond types 1-4) laid out like tasks/qm9_task.py:114-147 with
    tie_fwd_bkwd_edges=True: both directions of a bond live in the bond's type, adjacency sorted by (src, dst)."""
    n = int(np.clip(round(rng.normal(18.0, 3.0)), 3, 29))
    edges = []
    for v in range(1, n):                                   # random spanning tree
        edges.append((int(rng.integers(0, v)), v))
    for _ in range(max(0, int(round(0.035 * n + rng.random())))):   # a few ring closures
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        if a != b:
            edges.append((a, b))
    num_types = 4 + (1 if add_self_loop_edges else 0)
    shift = 1 if add_self_loop_edges else 0
    per_type = [[] for _ in range(num_types)]
    if add_self_loop_edges:
        per_type[0] = [(i, i) for i in range(n)]
    for (a, b) in edges:
        t = int(rng.choice(4, p=[0.86, 0.09, 0.03, 0.02])) + shift
        per_type[t].append((a, b))
        per_type[t].append((b, a))
    adj = []
    for lst in per_type:
        arr = np.array(sorted(lst), dtype=np.int32).reshape(-1, 2)
        adj.append(arr)
    feats = rng.standard_normal((n, feature_dim)).astype(np.float32)
    return GraphSample(adj, _in_degrees(adj, n), feats)


def make_qm9_like_graphs(num_graphs: int, seed: int = 0, add_self_loop_edges: bool = False) -> List[GraphSample]:
    rng = np.random.default_rng(seed)
    return [make_qm9_like_graph(rng, add_self_loop_edges=add_self_loop_edges) for _ in range(num_graphs)]


# ---- PPI files (tasks/ppi_task.py:68-160; the dgl "ppi.zip" layout: <fold>_graph.json, _feats.npy, _labels.npy, _graph_id.npy) ----
def load_ppi_fold(data_dir: str, fold: str = "train", add_self_loop_edges: bool = True, tie_fwd_bkwd_edges: bool = False):
    """Read one PPI data fold the way PPI_Task.__load_data does and return (graphs, labels): one GraphSample per graph id in
    order of first appearance, node ids shifted so every graph starts at 0 (:115-121,136-141), edge types in the reference's
    order -- 0 = forward links in file order, then the self-loop type if enabled (one (i, i) per node, in-degree 1), then the
    backward type (tgt, src) unless directions are tied (:99-106).  ``labels``: per graph a float32 [V_g, num_labels] array.
    (The data set itself is not shipped with the reference; the format is the public dgl download named at :69.)"""
    import json
    import os
    if fold not in ("train", "valid", "test"):
        raise ValueError("Unknown data fold '%s'" % str(fold))
    with open(os.path.join(data_dir, "%s_graph.json" % fold)) as f:
        links = json.load(f)["links"]
    feats = np.load(os.path.join(data_dir, "%s_feats.npy" % fold))
    labels = np.load(os.path.join(data_dir, "%s_labels.npy" % fold))
    graph_id = np.load(os.path.join(data_dir, "%s_graph_id.npy" % fold))
    num_types = 1 + (1 if add_self_loop_edges else 0) + (0 if tie_fwd_bkwd_edges else 1)
    self_type = 1 if add_self_loop_edges else None
    bkwd_type = None if tie_fwd_bkwd_edges else num_types - 1
    # graphs in order of first appearance; offset = id of the first node of the graph
    order, first = [], {}
    for node, g in enumerate(graph_id.tolist()):
        if g not in first:
            first[g] = node
            order.append(g)
    nodes_of = {g: np.nonzero(graph_id == g)[0] for g in order}
    src = np.asarray([e["source"] for e in links], dtype=np.int64)
    tgt = np.asarray([e["target"] for e in links], dtype=np.int64)
    link_graph = graph_id[src] if len(links) else np.zeros(0, dtype=graph_id.dtype)
    graphs, graph_labels = [], []
    for g in order:
        ids = nodes_of[g]
        n, off = int(ids.shape[0]), first[g]
        sel = link_graph == g                                   # links are assigned to the graph of their SOURCE (:137)
        fwd = np.stack([src[sel] - off, tgt[sel] - off], axis=1).astype(np.int32).reshape(-1, 2)
        adj = [None] * num_types
        adj[0] = fwd
        if self_type is not None:
            adj[self_type] = np.stack([np.arange(n), np.arange(n)], axis=1).astype(np.int32)
        if bkwd_type is not None:
            adj[bkwd_type] = fwd[:, ::-1].copy()
        graphs.append(GraphSample(adj, _in_degrees(adj, n), feats[ids].astype(np.float32)))
        graph_labels.append(labels[ids].astype(np.float32))
    return graphs, graph_labels


# ---- real QM9 records (data/qm9/*.jsonl.gz of the reference; tasks/qm9_task.py:85-147) ----
def qm9_num_edge_types(raw_graphs: Sequence[Dict], add_self_loop_edges: bool = True, tie_fwd_bkwd_edges: bool = True) -> int:
    """tasks/qm9_task.py:89-96: max bond type (+1 for the self-loop type 0), doubled when directions are untied."""
    num_fwd = max(max(e[1] for e in g["graph"]) for g in raw_graphs if len(g["graph"]))
    if add_self_loop_edges:
        num_fwd += 1
    return num_fwd * (1 if tie_fwd_bkwd_edges else 2)


def qm9_graph_to_sample(raw: Dict, num_edge_types: int, add_self_loop_edges: bool = True,
                        tie_fwd_bkwd_edges: bool = True) -> GraphSample:
    """One record {"graph": [(src, bond, dst)...], "node_features": [[15 floats]...]} -> per-type adjacency lists and
    in-degrees exactly as __graph_to_adjacency_lists builds them (tasks/qm9_task.py:114-147): bond types start at 1
    (0 is the self-loop type when enabled, else types are shifted down by one); tied directions put (dst, src) in the
    same type; each list is sorted by (src, dst); untied directions append the reversed lists as extra types.
    (tie_fwd_bkwd_edges=False cannot actually run in the reference: :139-145 appends to the list it enumerates and raises
    IndexError on the first molecule -- tests/test_reference_batcher_pin.py.  What is built here is what that loop evidently
    means, in-degree quirk included; every configuration the reference CAN run is pinned bit-exact by the same test.)"""
    num_nodes = len(raw["node_features"])
    lists: List[List] = [[] for _ in range(num_edge_types)]
    indeg = np.zeros((num_edge_types, num_nodes), dtype=np.float64)
    for src, e, dst in raw["graph"]:
        t = e if add_self_loop_edges else e - 1
        lists[t].append((src, dst))
        indeg[t, dst] += 1
        if tie_fwd_bkwd_edges:
            lists[t].append((dst, src))
            indeg[t, src] += 1
    if add_self_loop_edges:
        for v in range(num_nodes):
            indeg[0, v] = 1
            lists[0].append((v, v))
    adj = [np.array(sorted(l), dtype=np.int32).reshape(-1, 2) for l in lists]
    if not tie_fwd_bkwd_edges:
        half = num_edge_types // 2
        adj = adj[:half]
        for t in range(half):
            adj.append(np.array(sorted((y, x) for (x, y) in adj[t]), dtype=np.int32).reshape(-1, 2))
            for (x, y) in adj[t]:
                indeg[half + t][y] += 1      # as the reference counts it (:143-144): at the forward edge's target y,
                                             # although the reversed edge (y, x) arrives at x -- kept for identical feeds
    return GraphSample(adj, indeg.astype(np.float32), np.asarray(raw["node_features"], dtype=np.float32))


def load_qm9_jsonl(path: str, limit: Optional[int] = None) -> List[Dict]:
    """Read records of a reference data/qm9/*.jsonl.gz file (dpu_utils RichPath.read_by_file_suffix in the reference)."""
    import gzip
    import json
    out = []
    with gzip.open(path, "rt") as f:
        for line in f:
            out.append(json.loads(line))
            if limit is not None and len(out) >= limit:
                break
    return out


def qm9_records_from_structure(path: str, feature_dim: int = 15, seed: int = 0, num_targets: int = 13) -> List[Dict]:
    """Records in the layout of data/qm9/*.jsonl.gz rebuilt from a structure-only archive (tests/golden/make_qm9_structure.py:
    atoms per molecule + bonds of the reference's 10,000 validation molecules).  The graph structure is the real one; node
    features are seeded random numbers and targets are zero -- for timing the real batch shape, not for accuracy work."""
    z = np.load(path)
    sizes, nbonds, bonds = z["num_atoms"].astype(np.int64), z["num_bonds"].astype(np.int64), z["bonds"].astype(np.int64)
    rng = np.random.default_rng(seed)
    recs, b0 = [], 0
    for i in range(sizes.shape[0]):
        n, nb = int(sizes[i]), int(nbonds[i])
        recs.append({"id": "qm9-structure:%d" % i, "graph": bonds[b0:b0 + nb].tolist(),
                     "node_features": rng.standard_normal((n, feature_dim)).astype(np.float32),
                     "targets": [[0.0]] * num_targets})
        b0 += nb
    return recs


def qm9_batch(raw_graphs: Sequence[Dict], add_self_loop_edges: bool = True, tie_fwd_bkwd_edges: bool = True,
              task_ids: Sequence[int] = (0,), max_nodes_per_batch: Optional[int] = None):
    """Records -> (Batch, graph_nodes_list int32 [V], target_values float32 [len(task_ids), G]): the feed_dict of
    tasks/qm9_task.py:200-261 for one minibatch."""
    L = qm9_num_edge_types(raw_graphs, add_self_loop_edges, tie_fwd_bkwd_edges)
    samples = [qm9_graph_to_sample(g, L, add_self_loop_edges, tie_fwd_bkwd_edges) for g in raw_graphs]
    batch = pack_batch(samples, max_nodes_per_batch)
    G = batch.num_graphs
    sizes = np.diff(batch.graph_node_offsets)
    graph_nodes_list = np.repeat(np.arange(G, dtype=np.int32), sizes)
    targets = np.array([[raw_graphs[g]["targets"][t][0] for g in range(G)] for t in task_ids], dtype=np.float32)
    return batch, graph_nodes_list, targets


def make_typed_random_graph(num_nodes: int, num_edges: int, type_fractions: Sequence[float], feature_dim: int,
                            seed: int = 0) -> GraphSample:
    """Uniform random multigraph with the edges split over L types in the given proportions
    (VarMisuse-shaped config: SURVEY.md 8d config 5)."""
    rng = np.random.default_rng(seed)
    fr = np.asarray(type_fractions, dtype=np.float64)
    counts = np.floor(fr / fr.sum() * num_edges).astype(np.int64)
    counts[0] += num_edges - counts.sum()
    adj = []
    for c in counts:
        a = np.stack([rng.integers(0, num_nodes, size=int(c)), rng.integers(0, num_nodes, size=int(c))], axis=1)
        adj.append(a.astype(np.int32))
    feats = rng.standard_normal((num_nodes, feature_dim)).astype(np.float32)
    return GraphSample(adj, _in_degrees(adj, num_nodes), feats)


def pack_batch(graphs: Sequence[GraphSample], max_nodes_per_batch: Optional[int] = None) -> Batch:
    """The packing loop of tasks/ppi_task.py:213-251: add graphs while node_offset + |graph| <
    max_nodes_per_batch, shift node ids by the running offset, concatenate per type; an edge type with no
    edge in the batch becomes np.zeros((0, 2)) (:246-249)."""
    num_types = len(graphs[0].adjacency_lists)
    feats, indeg, offsets = [], [], [0]
    adj: List[List[np.ndarray]] = [[] for _ in range(num_types)]
    node_offset = 0
    for g in graphs:
        n = g.node_features.shape[0]
        if max_nodes_per_batch is not None and not (node_offset + n < max_nodes_per_batch):
            break
        feats.append(g.node_features)
        for i in range(num_types):
            adj[i].append(g.adjacency_lists[i].reshape(-1, 2) + node_offset)
        indeg.append(g.type_to_node_to_num_incoming_edges)
        node_offset += n
        offsets.append(node_offset)
    merged, num_edges = [], 0
    for i in range(num_types):
        a = np.concatenate(adj[i]).astype(np.int32) if len(adj[i]) > 0 else np.zeros((0, 2), dtype=np.int32)
        num_edges += a.shape[0]
        merged.append(a)
    return Batch(node_features=np.concatenate(feats, axis=0).astype(np.float32),
                 adjacency_lists=merged,
                 type_to_num_incoming_edges=np.concatenate(indeg, axis=1).astype(np.float32),
                 num_graphs=len(feats), num_nodes=node_offset, num_edges=num_edges,
                 graph_node_offsets=np.asarray(offsets, dtype=np.int64))


def minibatches(graphs: Sequence[GraphSample], max_nodes_per_batch: int) -> Iterator[Tuple[Batch, int]]:
    """One epoch of minibatches, the outer loop of tasks/ppi_task.py:211-256 (same in qm9_task.py:212-261): pack graphs in
    order until the next one would reach the node budget, emit, continue with that graph.  Yields (batch, index of its first
    graph).  A graph with >= max_nodes_per_batch nodes can never be packed -- the reference then spins on an empty batch
    (np.concatenate of an empty list raises); here it is a ValueError up front."""
    start = 0
    while start < len(graphs):
        n = graphs[start].node_features.shape[0]
        if not (n < max_nodes_per_batch):
            raise ValueError("graph %d has %d nodes: does not fit max_nodes_per_batch=%d" % (start, n, max_nodes_per_batch))
        batch = pack_batch(graphs[start:], max_nodes_per_batch)
        yield batch, start
        start += batch.num_graphs


def ppi_like_batch(num_graphs: int = 1, num_nodes: int = 2245, num_links: int = 59000, seed: int = 0,
                   zipf_targets: bool = False) -> Batch:
    """BASELINE config 2: one PPI-shaped graph -> V=2,245, M = 2*59,000 + 2,245 = 120,245, L=3."""
    return pack_batch([make_ppi_like_graph(num_nodes, num_links, seed=seed + i, zipf_targets=zipf_targets)
                       for i in range(num_graphs)])


def qm9_like_batch(num_graphs: int = 10000, seed: int = 0, add_self_loop_edges: bool = False) -> Batch:
    """BASELINE config 3 shape: 10k molecule graphs, ~18 nodes each, 4 bond types (5 with self loops)."""
    return pack_batch(make_qm9_like_graphs(num_graphs, seed, add_self_loop_edges))


VARMISUSE_TYPE_FRACTIONS = (0.30, 0.30, 0.15, 0.15, 0.05, 0.05)


def varmisuse_like_batch(num_nodes: int = 50000, num_edges: int = 1000000, packed_graphs: int = 0, seed: int = 0,
                         feature_dim: int = 64) -> Batch:
    """BASELINE config 5 shape: V=50k, M=1M, L=6.  packed_graphs=0 -> one random graph (worst-case halo);
    packed_graphs=g -> g equal graphs packed block-diagonally (zero-halo partition possible)."""
    if packed_graphs <= 0:
        return pack_batch([make_typed_random_graph(num_nodes, num_edges, VARMISUSE_TYPE_FRACTIONS, feature_dim, seed)])
    n, e = num_nodes // packed_graphs, num_edges // packed_graphs
    return pack_batch([make_typed_random_graph(n, e, VARMISUSE_TYPE_FRACTIONS, feature_dim, seed + i)
                       for i in range(packed_graphs)])
