"""CPU, property-based (hypothesis): invariants of the host-side data structures on random graphs -- the batch packer, the
graph-boundary and node-range shardings, and the reference-snapshot variable sorting."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st   # noqa: E402

from tf_gnn_samples_b200 import batching, checkpoint   # noqa: E402
from tf_gnn_samples_b200.partition import NodeRangePartition, split_batch_by_graphs   # noqa: E402


@st.composite
def graphs(draw, max_graphs=5):
    n_graphs = draw(st.integers(1, max_graphs))
    n_types = draw(st.integers(1, 4))
    seed = draw(st.integers(0, 10_000))
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_graphs):
        n = int(rng.integers(1, 30))
        adj = []
        for _t in range(n_types):
            e = int(rng.integers(0, 60))
            adj.append(np.stack([rng.integers(0, n, e), rng.integers(0, n, e)], axis=1).astype(np.int32).reshape(-1, 2))
        indeg = np.stack([np.bincount(a[:, 1], minlength=n) for a in adj]).astype(np.int32)
        out.append(batching.GraphSample(adj, indeg, rng.standard_normal((n, 3)).astype(np.float32)))
    return out


@settings(max_examples=40, deadline=None)
@given(graphs())
def test_pack_batch_invariants(gs):
    b = batching.pack_batch(gs)
    assert b.num_graphs == len(gs) and b.num_nodes == sum(g.node_features.shape[0] for g in gs)
    assert b.num_edges == sum(a.shape[0] for g in gs for a in g.adjacency_lists)
    off = b.graph_node_offsets
    assert off[0] == 0 and off[-1] == b.num_nodes and np.all(np.diff(off) > 0)
    for l, a in enumerate(b.adjacency_lists):
        assert a.dtype == np.int32 and a.shape[1] == 2
        if a.shape[0]:
            assert a.min() >= 0 and a.max() < b.num_nodes
            gs_src = np.searchsorted(off, a[:, 0], side="right") - 1
            gs_tgt = np.searchsorted(off, a[:, 1], side="right") - 1
            assert np.array_equal(gs_src, gs_tgt)                         # block-diagonal: no edge crosses a graph boundary
        assert np.array_equal(b.type_to_num_incoming_edges[l], np.bincount(a[:, 1], minlength=b.num_nodes))


@settings(max_examples=40, deadline=None)
@given(graphs(), st.integers(1, 4))
def test_graph_boundary_shards_partition_the_batch(gs, parts):
    b = batching.pack_batch(gs)
    shards = split_batch_by_graphs(b, parts)
    assert len(shards) == parts
    assert sum(s.num_nodes for s in shards) == b.num_nodes and sum(s.num_edges for s in shards) == b.num_edges
    assert sum(s.num_graphs for s in shards) == b.num_graphs
    lo = 0
    for s in shards:
        np.testing.assert_array_equal(s.node_features, b.node_features[lo:lo + s.num_nodes])
        for a in s.adjacency_lists:
            if a.shape[0]:
                assert a.min() >= 0 and a.max() < s.num_nodes
        lo += s.num_nodes


@settings(max_examples=30, deadline=None)
@given(graphs(max_graphs=3), st.integers(1, 4))
def test_node_range_partition_owns_every_edge_once(gs, world):
    b = batching.pack_batch(gs)
    parts = [NodeRangePartition(b.adjacency_lists, b.type_to_num_incoming_edges, b.num_nodes, r, world) for r in range(world)]
    assert sum(p.n_own for p in parts) == b.num_nodes and sum(p.num_local_edges for p in parts) == b.num_edges
    for r, p in enumerate(parts):
        assert p.n_local == p.n_own + p.n_halo
        ids = np.concatenate([np.arange(p.lo, p.hi), p.halo_global])
        for l, la in enumerate(p.local_adjacency_lists):                   # local edges map back to global edges with owned targets
            if la.shape[0]:
                assert la[:, 1].max() < p.n_own
                g = np.stack([ids[la[:, 0]], ids[la[:, 1]]], axis=1)
                a = b.adjacency_lists[l]
                mine = a[(a[:, 1] >= p.lo) & (a[:, 1] < p.hi)]
                assert np.array_equal(g, mine)
        # what peers send to r is exactly r's halo, grouped by owner
        assert int(p.recv_counts.sum()) == p.n_halo
        for q, other in enumerate(parts):
            if q != r:
                assert int(other.send_counts[r]) == int(p.recv_counts[q])


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 4), st.integers(1, 5), st.booleans())
def test_snapshot_sorting_is_order_independent(num_layers, num_types, with_adam):
    rng = np.random.default_rng(num_layers * 10 + num_types)
    w = {}
    for l in range(num_layers):
        for t in range(num_types):
            w["graph_model/gnn_layer_%d/Edge_%d_Weight/kernel:0" % (l, t)] = rng.standard_normal((4, 4)).astype(np.float32)
            if with_adam:
                w["graph_model/gnn_layer_%d/Edge_%d_Weight/kernel/Adam:0" % (l, t)] = np.zeros((4, 4), np.float32)
    keys = list(w)
    rng.shuffle(keys)
    s = checkpoint.sort_variables({k: w[k] for k in keys})
    assert s["layer_indices"] == list(range(num_layers)) and not s["unused"] and not s["outside"]
    for l, layer in enumerate(s["layers"]):
        assert len(layer["edge_weights"]) == num_types
        for t in range(num_types):
            np.testing.assert_array_equal(layer["edge_weights"][t], w["graph_model/gnn_layer_%d/Edge_%d_Weight/kernel:0" % (l, t)])
