"""RGDCN (gnns/rgdcn.py): CPU -- the vectorised oracle against a literal per-edge loop; GPU -- the engine against the
oracle for the four variants (full / channel state x tied / untied kernels), all aggregations, several channel shapes."""
import numpy as np
import pytest

from oracle import ref_layers as R
from tf_gnn_samples_b200 import weights as W

from helpers import assert_parity, node_states, tiny_graph


def loop_rgdcn(h, adj, indeg, C, K, full, tie, act, w):
    """One timestep, sum aggregation, written edge by edge (rgdcn.py:121-160)."""
    fn = R.get_activation(act)
    V = h.shape[0]
    out = np.zeros((V, C * K))
    for c in range(C):
        for l, a in enumerate(adj):
            F = np.asarray(w["channel_weights"][l][0 if tie else c], np.float64)
            for (u, v) in a:
                x = h[v] if full else h[v, c * K:(c + 1) * K]
                kern = R._apply_act(fn, x @ F).reshape(K, K)
                out[v, c * K:(c + 1) * K] += (h[u, c * K:(c + 1) * K] @ kern) / (indeg[l][v] + 1e-7)
    return R._apply_act(fn, out)


@pytest.mark.parametrize("full,tie", [(False, False), (False, True), (True, False), (True, True)])
def test_oracle_matches_a_per_edge_loop(full, tie):
    adj, indeg = tiny_graph(19, (30, 0, 21), seed=41)
    C, K = 3, 4
    h = node_states(19, C * K, seed=42).astype(np.float64)
    w = W.rgdcn_weights(3, C, K, full, tie, seed=43, stddev=0.4)
    got = R.sparse_rgdcn_layer(h, adj, indeg, C, K, 1, full, tie, "tanh", "sum", True, weights=w)
    want = loop_rgdcn(h, adj, indeg.astype(np.float64), C, K, full, tie, "tanh", w)
    assert R.max_norm_rel_err(got, want) < 1e-12


def test_oracle_default_initialiser_scale():
    """truncated_normal(stddev = 1 / K^2) (rgdcn.py:101): values bounded by 2 stddev."""
    w = W.rgdcn_weights(2, 8, 16)
    k = w["channel_weights"][0][0]
    assert k.shape == (16, 256) and np.abs(k).max() <= 2.0 / 256 + 1e-9 and len(w["channel_weights"][0]) == 8


CASES = [  # C, K, full, tie, act, agg, normalize, T
    (8, 16, False, False, "ReLU", "sum", True, 1),      # RGDCN_Model defaults (models/rgdcn_model.py:13-24)
    (8, 16, True, False, "tanh", "sum", True, 2),
    (8, 16, False, True, "tanh", "mean", False, 1),
    (8, 16, True, True, "gelu", "sqrt_n", True, 1),
    (4, 8, False, False, "tanh", "max", True, 2),
    (8, 32, False, False, "elu", "sum", True, 1),
    (16, 4, True, False, "tanh", "max", False, 1),
    (3, 64, False, False, "tanh", "sum", True, 1),      # D = 192: the last 128-column block is half used
]


@pytest.mark.gpu
@pytest.mark.parametrize("C,K,full,tie,act,agg,normalize,T", CASES)
def test_engine_matches_oracle(cuda_device, C, K, full, tie, act, agg, normalize, T):
    import torch
    from tf_gnn_samples_b200 import sparse_rgdcn_layer
    adj, indeg = tiny_graph(61, (170, 61, 0, 95), seed=44)
    h = node_states(61, C * K, seed=45)
    w = W.rgdcn_weights(4, C, K, full, tie, seed=46, stddev=0.5 / K)
    want = R.sparse_rgdcn_layer(h, adj, indeg, C, K, T, full, tie, act, agg, normalize, weights=w)
    got = sparse_rgdcn_layer(torch.as_tensor(h).to(cuda_device), adj, torch.as_tensor(indeg).to(cuda_device), C, K, T, full, tie,
                             act, agg, normalize, weights=W.to_torch(w, cuda_device))
    err = assert_parity(got.cpu().numpy(), want, "rgdcn C=%d K=%d full=%s tie=%s %s %s" % (C, K, full, tie, act, agg))
    print("rgdcn", C, K, full, tie, act, agg, "err %.2e" % err)


@pytest.mark.gpu
def test_engine_on_a_ppi_shaped_batch_and_errors(cuda_device):
    import torch
    from tf_gnn_samples_b200 import GraphPlan, RgnnError, batching, sparse_rgdcn_layer
    b = batching.ppi_like_batch(num_nodes=900, num_links=20000, seed=47)
    h = node_states(b.num_nodes, 128, seed=48)
    w = W.rgdcn_weights(3, 8, 16, seed=49, stddev=0.05)
    want = R.sparse_rgdcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, 8, 16, 1, activation_function="ReLU", weights=w)
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)
    hd, wd = torch.as_tensor(h).to(cuda_device), W.to_torch(w, cuda_device)
    got = sparse_rgdcn_layer(hd, plan, cnt, 8, 16, activation_function="ReLU", weights=wd)
    assert_parity(got.cpu().numpy(), want, "rgdcn PPI-shaped")
    with pytest.raises(RgnnError):
        sparse_rgdcn_layer(hd, plan, cnt, 8, 12, weights=wd)                       # 8 * 12 != 128
    with pytest.raises(ValueError):
        sparse_rgdcn_layer(hd, plan, cnt, 8, 16, activation_function="swish", weights=wd)
    with pytest.raises(RgnnError):
        sparse_rgdcn_layer(hd.requires_grad_(True), plan, cnt, 8, 16, weights=wd)   # no gradient path
