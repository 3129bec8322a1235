"""CPU: pin the analytic RGCN gradients of the oracle against central finite differences of the float64 forward."""
import numpy as np
import pytest

from oracle import ref_grads as RG
from oracle import ref_layers as R
from tf_gnn_samples_b200 import weights as W

from helpers import node_states, tiny_graph


@pytest.mark.parametrize("act,agg,normalize", [("tanh", "sum", True), ("ReLU", "mean", False), ("elu", "sqrt_n", True),
                                               ("gelu", "sum", True), ("linear", "sum", False), ("selu", "mean", True),
                                               ("leaky_relu", "sum", True)])
def test_rgcn_grads_match_finite_differences(act, agg, normalize):
    adj, indeg = tiny_graph(13, (30, 13, 0, 17), seed=41)
    D = 6
    h = node_states(13, D, seed=42).astype(np.float64) + 0.05      # keep ReLU kinks away from exact zeros
    w = {"edge_weights": [k.astype(np.float64) for k in W.rgcn_weights(4, D, D, seed=43)["edge_weights"]]}
    g = np.random.default_rng(44).standard_normal((13, D))

    def loss(hh, ww):
        return float(np.sum(g * R.sparse_rgcn_layer(hh, adj, indeg, D, activation_function=act,
                                                    message_aggregation_function=agg,
                                                    normalize_by_num_incoming=normalize, weights=ww)))

    d_h, d_ws = RG.rgcn_layer_grads(h, adj, indeg, g, act, agg, normalize, weights=w)
    eps = 1e-6
    rng = np.random.default_rng(45)
    for _ in range(25):                                             # random coordinates of h
        i, j = rng.integers(0, 13), rng.integers(0, D)
        hp, hm = h.copy(), h.copy(); hp[i, j] += eps; hm[i, j] -= eps
        fd = (loss(hp, w) - loss(hm, w)) / (2 * eps)
        assert abs(fd - d_h[i, j]) <= 1e-5 * max(1.0, abs(fd)), (i, j, fd, d_h[i, j])
    for _ in range(25):                                             # random coordinates of the kernels
        l, i, j = rng.integers(0, 4), rng.integers(0, D), rng.integers(0, D)
        wp = {"edge_weights": [k.copy() for k in w["edge_weights"]]}; wm = {"edge_weights": [k.copy() for k in w["edge_weights"]]}
        wp["edge_weights"][l][i, j] += eps; wm["edge_weights"][l][i, j] -= eps
        fd = (loss(h, wp) - loss(h, wm)) / (2 * eps)
        assert abs(fd - d_ws[l][i, j]) <= 1e-5 * max(1.0, abs(fd)), (l, i, j, fd, d_ws[l][i, j])
    assert np.all(d_ws[2] == 0)                                     # edge type with no edges gets a zero gradient
