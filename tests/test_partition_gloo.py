"""CPU, world_size=2 over gloo: the multi-GPU sharding logic.  The layer arithmetic on the
CPU side of these tests is the ORACLE (tests may use it as the checker); the product's compute has no CPU
path.  What is verified: (1) graph-boundary shards are independent and reassemble to the batch result,
(2) the node-range partition + all-to-all-v halo exchange hands every rank exactly the source rows it needs,
so that running a layer on the rank-local graph reproduces the global result on the owned rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_layers as R
from tf_gnn_samples_b200 import batching, weights as W
from tf_gnn_samples_b200.partition import NodeRangePartition, balanced_cuts, split_batch_by_graphs


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_balanced_cuts():
    cuts = balanced_cuts(np.array([5, 1, 1, 1, 1, 1, 5, 5]), 2)
    assert cuts[0] == 0 and cuts[-1] == 8 and 0 < cuts[1] < 8
    left = np.array([5, 1, 1, 1, 1, 1, 5, 5])[:cuts[1]].sum()
    assert abs(left - 10) <= 5


def test_graph_boundary_shards_are_independent():
    gs = [batching.make_ppi_like_graph(60 + 7 * i, 400, seed=i) for i in range(5)]
    b = batching.pack_batch(gs)
    D = 16
    h = np.tanh(np.random.default_rng(0).standard_normal((b.num_nodes, D)))
    w = W.rgcn_weights(3, D, D)
    full = R.sparse_rgcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, D, activation_function="relu", weights=w)
    shards = split_batch_by_graphs(b, 3)
    assert sum(s.num_nodes for s in shards) == b.num_nodes and sum(s.num_edges for s in shards) == b.num_edges
    outs, lo = [], 0
    for s in shards:
        outs.append(R.sparse_rgcn_layer(h[lo:lo + s.num_nodes], s.adjacency_lists, s.type_to_num_incoming_edges, D,
                                        activation_function="relu", weights=w))
        lo += s.num_nodes
    np.testing.assert_allclose(np.concatenate(outs), full, rtol=1e-12, atol=1e-12)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = batching.varmisuse_like_batch(num_nodes=600, num_edges=9000, seed=3, feature_dim=8)   # one random graph
        D = 16
        h = np.tanh(np.random.default_rng(1).standard_normal((b.num_nodes, D))).astype(np.float32)
        part = NodeRangePartition(b.adjacency_lists, b.type_to_num_incoming_edges, b.num_nodes, rank, world)
        h_own = torch.as_tensor(h[part.lo:part.hi])
        local = part.exchange(h_own)                                   # all-to-all-v over gloo
        ids = np.concatenate([np.arange(part.lo, part.hi), part.halo_global])
        ok_rows = bool(np.array_equal(local.numpy(), h[ids]))
        # run FiLM (target-dependent modulation + layer norm) on the rank-local graph with the oracle
        w = W.film_weights(len(b.adjacency_lists), D, D, random_ln=True)
        loc = R.sparse_gnn_film_layer(local.numpy(), part.local_adjacency_lists, part.local_num_incoming, D,
                                      normalize_by_num_incoming=True, weights=w)[:part.n_own]
        glob = R.sparse_gnn_film_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, D,
                                       normalize_by_num_incoming=True, weights=w)[part.lo:part.hi]
        err = R.max_norm_rel_err(loc, glob)
        edges = torch.tensor([part.num_local_edges], dtype=torch.int64)
        dist.all_reduce(edges)
        ret[rank] = (ok_rows, err, int(edges.item()), b.num_edges, part.n_halo, part.n_own)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_node_range_partition_halo_exchange_world2():
    world, port = 2, free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    assert set(res) == {0, 1}
    for rank, (ok_rows, err, edges_total, edges_batch, n_halo, n_own) in res.items():
        assert ok_rows, "rank %d received wrong halo rows" % rank
        assert err < 1e-9, "rank %d local result differs from global: %g" % (rank, err)
        assert edges_total == edges_batch                                # every edge owned by exactly one rank
        assert n_halo > 0 and n_own > 0


def _ddp_worker(rank, world, port, ret):
    """Data-parallel training host logic: shard the batch at graph boundaries, scale the local loss by the GLOBAL node
    count, sum gradients with scaffold.all_reduce_gradients_ -> must equal the gradient on the union batch.  The layer
    arithmetic is the float64 autograd oracle (CPU); the product's compute has no CPU path."""
    from oracle import ref_autograd as A
    from tf_gnn_samples_b200.scaffold import all_reduce_gradients_, global_count
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gs = [batching.make_ppi_like_graph(50 + 9 * i, 300, feature_dim=8, seed=20 + i) for i in range(4)]
        full = batching.pack_batch(gs)
        D = 8
        w_np = W.rgcn_weights(3, D, D, seed=5)
        labels_full = np.random.default_rng(6).standard_normal((full.num_nodes, D))

        def loss_total(b, labels, weights):
            out = A.sparse_rgcn_layer(torch.as_tensor(b.node_features, dtype=torch.float64), b.adjacency_lists,
                                      torch.as_tensor(b.type_to_num_incoming_edges, dtype=torch.float64), weights=weights)
            return ((out - torch.as_tensor(labels)) ** 2).sum()

        # union-batch gradient (what one device computes): d (total / V) / d W
        w_ref = A.to_torch64(w_np)
        (loss_total(full, labels_full, w_ref) / full.num_nodes).backward()
        # this rank's shard
        shards = split_batch_by_graphs(full, world)
        lo = sum(s.num_nodes for s in shards[:rank])
        mine = shards[rank]
        w_loc = A.to_torch64(w_np)
        params = w_loc["edge_weights"] + [torch.zeros(3, dtype=torch.float64, requires_grad=True)]   # one parameter never touched locally
        v_total = global_count(mine.num_nodes, torch.device("cpu"))
        (loss_total(mine, labels_full[lo:lo + mine.num_nodes], w_loc) / v_total).backward()
        all_reduce_gradients_(params)
        err = max(float((p.grad - q.grad).abs().max() / q.grad.abs().max()) for p, q in zip(w_loc["edge_weights"], w_ref["edge_weights"]))
        ret[rank] = (v_total, full.num_nodes, err, float(params[-1].grad.abs().max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_data_parallel_gradient_sync_world2():
    world, port = 2, free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    assert set(res) == {0, 1}
    for rank, (v_total, v_full, err, unused_grad) in res.items():
        assert v_total == v_full
        assert err < 1e-12, "rank %d: summed shard gradients differ from the union-batch gradient: %g" % (rank, err)
        assert unused_grad == 0.0


def _halo_train_worker(rank, world, port, ret):
    """Training ONE graph over a node-range partition: forward halo exchange, layer on the rank-local graph, backward through
    the transposed exchange (halo-row gradients return to their owners and are summed), weight gradients all-reduced.
    Everything must equal single-process autograd on the whole graph.  Layer arithmetic: the float64 autograd oracle."""
    from oracle import ref_autograd as A
    from tf_gnn_samples_b200.scaffold import all_reduce_gradients_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = batching.varmisuse_like_batch(num_nodes=300, num_edges=4000, seed=9, feature_dim=8)     # one random graph
        D, Ltypes = 8, len(b.adjacency_lists)
        h = np.tanh(np.random.default_rng(2).standard_normal((b.num_nodes, D)))
        proj = np.random.default_rng(3).standard_normal((b.num_nodes, D))
        w_np = W.film_weights(Ltypes, D, D, seed=4, random_ln=True)

        def two_layers(states, adj, cnt, weights, exchange=None, n_own=None):
            cur = states
            for _ in range(2):                                           # same weights twice: gradients accumulate
                local = exchange(cur) if exchange is not None else cur
                out = A.sparse_gnn_film_layer(local, adj, cnt, activation_function="tanh", normalize_by_num_incoming=True, weights=weights)
                cur = out[:n_own] if n_own is not None else out
            return cur

        # single process, whole graph
        h_ref = torch.as_tensor(h).requires_grad_(True)
        w_ref = A.to_torch64(w_np)
        (two_layers(h_ref, b.adjacency_lists, torch.as_tensor(b.type_to_num_incoming_edges, dtype=torch.float64), w_ref) * torch.as_tensor(proj)).sum().backward()
        # this rank's share
        part = NodeRangePartition(b.adjacency_lists, b.type_to_num_incoming_edges, b.num_nodes, rank, world)
        h_own = torch.as_tensor(h[part.lo:part.hi]).requires_grad_(True)
        w_loc = A.to_torch64(w_np)
        out = two_layers(h_own, part.local_adjacency_lists, torch.as_tensor(part.local_num_incoming, dtype=torch.float64), w_loc,
                         exchange=part.exchange, n_own=part.n_own)
        (out * torch.as_tensor(proj[part.lo:part.hi])).sum().backward()
        params = list(A.flatten(w_loc).values())
        all_reduce_gradients_(params)
        err_h = float((h_own.grad - h_ref.grad[part.lo:part.hi]).abs().max() / h_ref.grad.abs().max())
        err_w = max(float((p.grad - q.grad).abs().max() / max(float(q.grad.abs().max()), 1e-30))
                    for p, q in zip(params, A.flatten(w_ref).values()))
        ret[rank] = (err_h, err_w, part.n_halo)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_training_through_the_halo_exchange_world2():
    world, port = 2, free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_halo_train_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    assert set(res) == {0, 1}
    for rank, (err_h, err_w, n_halo) in res.items():
        assert n_halo > 0
        assert err_h < 1e-10, "rank %d: d/d node states through the halo exchange differs: %g" % (rank, err_h)
        assert err_w < 1e-10, "rank %d: all-reduced weight gradients differ: %g" % (rank, err_w)


def _no_halo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # two equal graphs packed block-diagonally: the degree-balanced cut falls on the graph boundary -> no halo anywhere
        g = batching.make_typed_random_graph(40, 300, (0.5, 0.5), 4, seed=1)
        b = batching.pack_batch([g, g])
        part = NodeRangePartition(b.adjacency_lists, b.type_to_num_incoming_edges, b.num_nodes, rank, world)
        calls = {"n": 0}
        real = dist.all_to_all_single

        def counting(*a, **k):
            calls["n"] += 1
            return real(*a, **k)
        dist.all_to_all_single = counting
        try:
            h_own = torch.randn(part.n_own, 4, dtype=torch.float64, requires_grad=True)
            local = part.exchange(h_own)
            local.sum().backward()
        finally:
            dist.all_to_all_single = real
        ret[rank] = (part.n_halo, part.any_halo(), calls["n"], tuple(local.shape), float(h_own.grad.min()), (part.lo, part.hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_exchange_skips_the_collective_when_no_rank_has_a_halo_world2():
    world, port = 2, free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_no_halo_worker, args=(world, port, ret), nprocs=world, join=True)
        res = dict(ret)
    for rank, (n_halo, any_halo, calls, shape, gmin, rng) in res.items():
        assert rng in ((0, 40), (40, 80)), rng
        assert n_halo == 0 and any_halo is False and calls == 0
        assert shape == (40, 4) and gmin == 1.0
