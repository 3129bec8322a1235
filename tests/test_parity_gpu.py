"""GPU parity: every layer function of the CUDA engine (through the ctypes C ABI) against the numpy
float64 oracle on identical seeded inputs.  Bar: max-norm relative error <= 1e-4 (north star)."""
import numpy as np
import pytest

from oracle import ref_layers as R
from tf_gnn_samples_b200 import (GraphPlan, RgnnError, batching, weights as W, sparse_ggnn_layer,
                                 sparse_gnn_edge_mlp_layer, sparse_gnn_film_layer, sparse_rgat_layer,
                                 sparse_rgcn_layer, sparse_rgin_layer)

from helpers import assert_parity, assert_parity_8c, node_states, tiny_graph, to_cuda_inputs

pytestmark = pytest.mark.gpu

ACTS = ["tanh", "ReLU", "leaky_relu", "elu", "selu", "gelu", "linear", None]
AGGS = ["sum", "max", "mean", "sqrt_n"]


def run(layer, ref, h, adj, indeg, device, w, with_indeg, **kw):
    import torch
    ht, adjt, ct = to_cuda_inputs(h, adj, indeg, device)
    wt = W.to_torch(w, device)
    if with_indeg:
        got = layer(ht, adjt, ct, **kw, weights=wt)
        want = ref(h, adj, indeg, **kw, weights=w)
    else:
        got = layer(ht, adjt, **kw, weights=wt)
        want = ref(h, adj, **kw, weights=w)
    torch.cuda.synchronize()
    return got.cpu().numpy(), want


def oracle32(ref, h, adj, indeg, w, with_indeg, **kw):
    """The reference op order in float32 (the stand-in for the TF1 CPU arithmetic): yardstick of SURVEY.md 8(c)."""
    return ref(h, adj, indeg, **kw, weights=w, dtype=np.float32) if with_indeg else ref(h, adj, **kw, weights=w, dtype=np.float32)


# ---------------------------------------------------------------- plan ---------------------------------
def test_plan_matches_stable_sort(cuda_device):
    adj, _ = tiny_graph(53, (200, 0, 77, 5), seed=3)
    plan = GraphPlan(adj, 53, device=cuda_device)
    ex = {k: v.cpu().numpy() for k, v in plan.export().items()}
    src = np.concatenate([a[:, 0] for a in adj]); tgt = np.concatenate([a[:, 1] for a in adj])
    typ = np.concatenate([np.full(a.shape[0], l) for l, a in enumerate(adj)])
    order = np.lexsort((np.arange(src.size), typ, tgt))           # (target, type, original position)
    assert plan.num_edges == src.size
    np.testing.assert_array_equal(ex["e_orig"], order)
    np.testing.assert_array_equal(ex["e_src"], src[order])
    np.testing.assert_array_equal(ex["e_type"], typ[order])
    np.testing.assert_array_equal(ex["seg_off"], np.concatenate([[0], np.cumsum(np.bincount(tgt, minlength=53))]))


def test_plan_rejects_out_of_range_ids(cuda_device):
    bad = [np.array([[0, 1], [2, 99]], dtype=np.int32)]
    with pytest.raises(RgnnError):
        GraphPlan(bad, 10, device=cuda_device)


def test_plan_all_types_empty(cuda_device):
    import torch
    adj = [np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32)]
    h = node_states(9, 8)
    w = W.rgcn_weights(2, 8, 8)
    got = sparse_rgcn_layer(torch.as_tensor(h).to(cuda_device), adj, np.zeros((2, 9), np.float32), 8,
                            activation_function="tanh", weights=W.to_torch(w, cuda_device))
    assert np.all(got.cpu().numpy() == 0.0)                      # act(0) for every node


# ---------------------------------------------------------------- RGCN ---------------------------------
@pytest.mark.parametrize("act", ACTS)
@pytest.mark.parametrize("agg", AGGS)
def test_rgcn_small(cuda_device, act, agg):
    adj, indeg = tiny_graph()
    h = node_states(37, 64)
    w = W.rgcn_weights(4, 64, 64)
    got, want = run(sparse_rgcn_layer, R.sparse_rgcn_layer, h, adj, indeg, cuda_device, w, True, state_dim=64,
                    activation_function=act, message_aggregation_function=agg)
    if agg == "max":   # empty segments hold float32 lowest() through the activation (A.2): compare finite part
        mask = np.abs(want) < 1e30
        assert np.array_equal(mask, np.abs(got) < 1e30)
        assert_parity(np.where(mask, got, 0), np.where(mask, want, 0), "rgcn %s %s" % (act, agg))
    else:
        assert_parity(got, want, "rgcn %s %s" % (act, agg))


@pytest.mark.parametrize("d_in,d_out", [(64, 128), (128, 64), (52, 36), (320, 320), (512, 512)])
def test_rgcn_dims(cuda_device, d_in, d_out):
    adj, indeg = tiny_graph(101, (300, 17, 0, 250), seed=5)
    h = node_states(101, d_in)
    w = W.rgcn_weights(4, d_in, d_out)
    got, want = run(sparse_rgcn_layer, R.sparse_rgcn_layer, h, adj, indeg, cuda_device, w, True, state_dim=d_out,
                    activation_function="relu")
    assert_parity(got, want, "rgcn dims %d->%d" % (d_in, d_out))


@pytest.mark.parametrize("normalize,both,T", [(True, False, 3), (False, False, 1), (True, True, 1), (False, True, 2)])
def test_rgcn_options(cuda_device, normalize, both, T):
    adj, indeg = tiny_graph(64, (150, 64, 150), seed=7)
    h = node_states(64, 96)
    w = W.rgcn_weights(3, 96, 96, use_both_source_and_target=both)
    got, want = run(sparse_rgcn_layer, R.sparse_rgcn_layer, h, adj, indeg, cuda_device, w, True, state_dim=96,
                    num_timesteps=T, activation_function="tanh", normalize_by_num_incoming=normalize,
                    use_both_source_and_target=both)
    assert_parity(got, want, "rgcn normalize=%s both=%s T=%d" % (normalize, both, T))


def test_rgcn_ppi_shaped_hidden256(cuda_device):
    """BASELINE config 2: V=2,245, M=120,245, L=3, D=256, ReLU, sum, normalised."""
    b = batching.ppi_like_batch()
    assert (b.num_nodes, b.num_edges) == (2245, 120245)
    h = node_states(b.num_nodes, 256)
    w = W.rgcn_weights(3, 256, 256)
    got, want = run(sparse_rgcn_layer, R.sparse_rgcn_layer, h, b.adjacency_lists, b.type_to_num_incoming_edges,
                    cuda_device, w, True, state_dim=256, activation_function="ReLU")
    err = assert_parity(got, want, "rgcn PPI-shaped")
    print("rgcn PPI-shaped max-norm rel err %.3e" % err)


def test_rgcn_zipf_skew_and_plan_reuse(cuda_device):
    import torch
    b = batching.ppi_like_batch(num_nodes=1500, num_links=40000, zipf_targets=True, seed=4)
    h = node_states(b.num_nodes, 128)
    ws = [W.rgcn_weights(3, 128, 128, seed=10 + i) for i in range(2)]
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device)
    cur = torch.as_tensor(h).to(cuda_device)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)
    ref = h
    for w in ws:   # two stacked layers sharing one plan
        cur = sparse_rgcn_layer(cur, plan, cnt, 128, activation_function="relu", weights=W.to_torch(w, cuda_device))
        ref = R.sparse_rgcn_layer(ref, b.adjacency_lists, b.type_to_num_incoming_edges, 128,
                                  activation_function="relu", weights=w)
    assert_parity(cur.cpu().numpy(), ref, "rgcn zipf 2 layers")


def test_rgcn_deterministic(cuda_device):
    import torch
    b = batching.ppi_like_batch(num_nodes=800, num_links=20000, seed=9)
    h = torch.as_tensor(node_states(800, 128)).to(cuda_device)
    w = W.to_torch(W.rgcn_weights(3, 128, 128), cuda_device)
    outs = []
    for _ in range(3):
        outs.append(sparse_rgcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, 128,
                                      activation_function="tanh", weights=w).cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[1], outs[2])   # sorted segments: bit-reproducible


# ---------------------------------------------------------------- GGNN ---------------------------------
@pytest.mark.parametrize("cell,act,T", [("gru", "tanh", 1), ("GRU", "tanh", 4), ("rnn", "tanh", 2), ("gru", "relu", 2)])
def test_ggnn(cuda_device, cell, act, T):
    b = batching.qm9_like_batch(40, seed=2, add_self_loop_edges=(T == 2))
    D = 128
    h = node_states(b.num_nodes, D)
    w = W.ggnn_weights(len(b.adjacency_lists), D, cell=cell, random_bias=True)
    got, want = run(sparse_ggnn_layer, R.sparse_ggnn_layer, h, b.adjacency_lists, None, cuda_device, w, False,
                    state_dim=D, num_timesteps=T, gated_unit_type=cell, activation_function=act)
    assert_parity(got, want, "ggnn %s %s T=%d" % (cell, act, T))


@pytest.mark.parametrize("agg", ["max", "mean"])
def test_ggnn_aggregations(cuda_device, agg):
    adj, _ = tiny_graph(45, (90, 45, 30), seed=11, with_isolated=False)
    for a in adj:   # every node gets an incoming edge so that 'max' has no lowest() rows feeding the GRU
        pass
    adj[1] = np.stack([np.arange(45), np.arange(45)], axis=1).astype(np.int32)
    h = node_states(45, 64)
    w = W.ggnn_weights(3, 64)
    got, want = run(sparse_ggnn_layer, R.sparse_ggnn_layer, h, adj, None, cuda_device, w, False, state_dim=64,
                    num_timesteps=2, message_aggregation_function=agg)
    assert_parity(got, want, "ggnn agg %s" % agg)


# ---------------------------------------------------------------- RGAT ---------------------------------
@pytest.mark.parametrize("D,K,T", [(64, 4, 1), (256, 8, 1), (128, 4, 2), (96, 1, 1), (64, 16, 1)])
def test_rgat(cuda_device, D, K, T):
    adj, _ = tiny_graph(71, (260, 71, 0, 140), seed=13)
    h = node_states(71, D)
    w = W.rgat_weights(4, D, D)
    got, want = run(sparse_rgat_layer, R.sparse_rgat_layer, h, adj, None, cuda_device, w, False, state_dim=D,
                    num_timesteps=T, num_heads=K, activation_function="tanh")
    assert_parity(got, want, "rgat D=%d K=%d T=%d" % (D, K, T))


def test_rgat_ppi_shaped(cuda_device):
    b = batching.ppi_like_batch(num_nodes=1200, num_links=30000, seed=3)
    h = node_states(b.num_nodes, 256)
    w = W.rgat_weights(3, 256, 256)
    got, want = run(sparse_rgat_layer, R.sparse_rgat_layer, h, b.adjacency_lists, None, cuda_device, w, False,
                    state_dim=256, num_heads=8, activation_function="tanh")
    assert_parity(got, want, "rgat PPI-shaped")


# ---------------------------------------------------------------- FiLM ---------------------------------
@pytest.mark.parametrize("normalize,agg,act,T", [(False, "sum", "ReLU", 1), (True, "sum", "tanh", 2),
                                                  (False, "mean", "gelu", 1), (False, "max", "elu", 1)])
def test_film(cuda_device, normalize, agg, act, T):
    adj, indeg = tiny_graph(83, (300, 83, 120, 0, 40), seed=17, with_isolated=(agg != "max"))
    if agg == "max":
        adj[1] = np.stack([np.arange(83), np.arange(83)], axis=1).astype(np.int32)
        indeg = np.stack([np.bincount(a[:, 1], minlength=83) for a in adj]).astype(np.float32)
    D = 128
    h = node_states(83, D)
    w = W.film_weights(5, D, D, num_timesteps=T, random_ln=True)
    kw = dict(state_dim=D, num_timesteps=T, activation_function=act, message_aggregation_function=agg,
              normalize_by_num_incoming=normalize)
    got, want = run(sparse_gnn_film_layer, R.sparse_gnn_film_layer, h, adj, indeg, cuda_device, w, True, **kw)
    assert_parity_8c(got, want, oracle32(R.sparse_gnn_film_layer, h, adj, indeg, w, True, **kw),
                     "film norm=%s %s %s T=%d" % (normalize, agg, act, T))


# ---------------------------------------------------------------- Edge-MLP -----------------------------
@pytest.mark.parametrize("hidden,use_target,normalize,T", [(0, True, False, 1), (1, True, False, 1), (2, True, True, 1),
                                                           (1, False, False, 1), (0, False, True, 2), (1, True, False, 2)])
def test_edge_mlp(cuda_device, hidden, use_target, normalize, T):
    adj, indeg = tiny_graph(67, (210, 67, 0, 95), seed=19)
    D = 64
    h = node_states(67, D)
    w = W.edge_mlp_weights(4, D, D, num_edge_hidden_layers=hidden, use_target_state_as_input=use_target,
                           num_timesteps=T, random_ln=True)
    kw = dict(state_dim=D, num_timesteps=T, activation_function="gelu", normalize_by_num_incoming=normalize,
              use_target_state_as_input=use_target, num_edge_hidden_layers=hidden)
    got, want = run(sparse_gnn_edge_mlp_layer, R.sparse_gnn_edge_mlp_layer, h, adj, indeg, cuda_device, w, True, **kw)
    assert_parity_8c(got, want, oracle32(R.sparse_gnn_edge_mlp_layer, h, adj, indeg, w, True, **kw),
                     "edge-mlp h=%d tgt=%s norm=%s T=%d" % (hidden, use_target, normalize, T))


# ---------------------------------------------------------------- RGIN ---------------------------------
@pytest.mark.parametrize("edge_h,aggr_h,use_target,T", [(1, None, False, 1), (1, 1, False, 1), (0, None, True, 1),
                                                        (2, 0, True, 1), (None, 1, False, 1), (None, 0, True, 1),
                                                        (1, None, False, 2)])
def test_rgin(cuda_device, edge_h, aggr_h, use_target, T):
    adj, _ = tiny_graph(59, (180, 59, 77), seed=23)
    D = 64
    h = node_states(59, D)
    w = W.rgin_weights(3, D, D, num_edge_MLP_hidden_layers=edge_h, num_aggr_MLP_hidden_layers=aggr_h,
                       use_target_state_as_input=use_target, num_timesteps=T, random_ln=True)
    kw = dict(state_dim=D, num_timesteps=T, activation_function="ReLU", use_target_state_as_input=use_target,
              num_edge_MLP_hidden_layers=edge_h, num_aggr_MLP_hidden_layers=aggr_h)
    got, want = run(sparse_rgin_layer, R.sparse_rgin_layer, h, adj, None, cuda_device, w, False, **kw)
    assert_parity_8c(got, want, oracle32(R.sparse_rgin_layer, h, adj, None, w, False, **kw),
                     "rgin edge=%s aggr=%s tgt=%s T=%d" % (edge_h, aggr_h, use_target, T))


# ---------------------------------------------------------------- errors --------------------------------
def test_error_behaviour(cuda_device):
    import torch
    adj, indeg = tiny_graph()
    h = torch.as_tensor(node_states(37, 64)).to(cuda_device)
    w = W.to_torch(W.rgcn_weights(4, 64, 64), cuda_device)
    with pytest.raises(ValueError, match="Unknown activation function"):
        sparse_rgcn_layer(h, adj, indeg, 64, activation_function="swish", weights=w)
    with pytest.raises(ValueError, match="Unknown aggregation function"):
        sparse_rgcn_layer(h, adj, indeg, 64, message_aggregation_function="median", weights=w)
    with pytest.raises(Exception, match="Unknown RNN cell type"):
        sparse_ggnn_layer(h, adj, 64, gated_unit_type="transformer", weights=W.to_torch(W.ggnn_weights(4, 64), cuda_device))
    with pytest.raises(RgnnError):   # no CPU path: CPU tensors are rejected loudly
        sparse_rgcn_layer(h.cpu(), adj, indeg, 64, weights=w)
    with pytest.raises(RgnnError):   # num_timesteps > 1 needs state_dim == D (rgcn.py docstring)
        sparse_rgcn_layer(h, adj, indeg, 32, num_timesteps=2, weights=W.to_torch(W.rgcn_weights(4, 64, 32), cuda_device))


# ---------------------------------------------------------------- static-weight cache --------------------
def test_weight_cache_tracks_in_place_updates(cuda_device):
    import torch
    import tf_gnn_samples_b200 as G
    b = batching.ppi_like_batch(num_nodes=500, num_links=6000, seed=5)
    h = node_states(b.num_nodes, 128)
    w = W.rgcn_weights(3, 128, 128, seed=77)
    wt = W.to_torch(w, cuda_device)
    ht = torch.as_tensor(h).to(cuda_device)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device)
    G.set_weight_cache(True)
    try:
        first = sparse_rgcn_layer(ht, plan, cnt, 128, activation_function="tanh", weights=wt).cpu().numpy()
        again = sparse_rgcn_layer(ht, plan, cnt, 128, activation_function="tanh", weights=wt).cpu().numpy()
        assert np.array_equal(first, again)                          # second call served from the cached images
        assert_parity(first, R.sparse_rgcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, 128,
                                                 activation_function="tanh", weights=w), "cached weights")
        for k in wt["edge_weights"]:
            k.mul_(0.5)                                              # in-place update bumps Tensor._version
        w2 = {"edge_weights": [0.5 * k for k in w["edge_weights"]]}
        changed = sparse_rgcn_layer(ht, plan, cnt, 128, activation_function="tanh", weights=wt).cpu().numpy()
        assert_parity(changed, R.sparse_rgcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, 128,
                                                   activation_function="tanh", weights=w2), "cache flushed on update")
    finally:
        G.set_weight_cache(False)


def test_rgcn_layer_stack_equals_layer_by_layer(cuda_device):
    import torch
    import tf_gnn_samples_b200 as G
    b = batching.ppi_like_batch(num_nodes=700, num_links=9000, seed=8)
    h = torch.as_tensor(node_states(b.num_nodes, 128)).to(cuda_device)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)
    ws = [W.to_torch(W.rgcn_weights(3, 128, 128, seed=20 + i), cuda_device) for i in range(3)]
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device, validate=False)   # deferred index check
    plan.check()
    cur = h
    for w in ws:
        cur = sparse_rgcn_layer(cur, plan, cnt, 128, activation_function="ReLU", weights=w)
    stacked = G.rgcn_layer_stack(h, plan, cnt, ws, activation_function="ReLU")
    assert torch.equal(cur, stacked)


def test_deferred_plan_check_reports_bad_ids(cuda_device):
    bad = [np.array([[0, 1], [2, 99]], dtype=np.int32)]
    plan = GraphPlan(bad, 10, device=cuda_device, validate=False)      # no exception yet: check is deferred
    with pytest.raises(RgnnError):
        plan.check()


# ---------------- degree skew: heavy targets are reduced by a whole CTA (seg_reduce_heavy_kernel) ----------------
@pytest.mark.parametrize("validate", [True, False])
def test_heavy_segments_rgcn_and_film(cuda_device, validate):
    import torch
    b = batching.ppi_like_batch(num_nodes=1200, num_links=60000, zipf_targets=True, seed=21)
    deg = b.type_to_num_incoming_edges.sum(axis=0)
    assert deg.max() > 2000                                        # far beyond the 512-edge split threshold
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device, validate=validate)
    D = 128
    h = node_states(b.num_nodes, D)
    ht = torch.as_tensor(h).to(cuda_device)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)
    for agg in ("sum", "max", "mean"):
        w = W.rgcn_weights(3, D, D, seed=31)
        got = sparse_rgcn_layer(ht, plan, cnt, D, activation_function="tanh", message_aggregation_function=agg,
                                weights=W.to_torch(w, cuda_device)).cpu().numpy()
        want = R.sparse_rgcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, D, activation_function="tanh",
                                   message_aggregation_function=agg, weights=w)
        assert_parity(got, want, "heavy rgcn %s validate=%s" % (agg, validate))
    w = W.film_weights(3, D, D, random_ln=True)
    got = sparse_gnn_film_layer(ht, plan, cnt, D, normalize_by_num_incoming=True, weights=W.to_torch(w, cuda_device)).cpu().numpy()
    want = R.sparse_gnn_film_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, D, normalize_by_num_incoming=True, weights=w)
    assert_parity(got, want, "heavy film validate=%s" % validate)
    plan.check()
