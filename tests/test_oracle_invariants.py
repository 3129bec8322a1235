"""CPU: structural invariants every layer of the oracle must satisfy (they follow from the reference's definitions: sums over
incoming edges, weights shared across nodes) -- edge-order permutation, node relabelling equivariance, isolated extra nodes.
The GPU parity tests compare against this oracle, so pinning its structure pins theirs."""
import numpy as np
import pytest

from oracle import ref_layers as R
from tf_gnn_samples_b200 import weights as W

from helpers import node_states, tiny_graph

V, D, L = 29, 8, 3


def cases():
    adj, indeg = tiny_graph(V, (50, 0, 33), seed=71, duplicates=True)
    h = node_states(V, D, seed=72).astype(np.float64)
    return {
        "rgcn": (lambda h, adj, c: R.sparse_rgcn_layer(h, adj, c, D, num_timesteps=2, weights=W.rgcn_weights(L, D, D, seed=1)), True),
        "rgcn_max_both": (lambda h, adj, c: R.sparse_rgcn_layer(h, adj, c, D, message_aggregation_function="max", use_both_source_and_target=True,
                                                               weights=W.rgcn_weights(L, D, D, seed=2, use_both_source_and_target=True)), True),
        "ggnn": (lambda h, adj, c: R.sparse_ggnn_layer(h, adj, D, num_timesteps=2, weights=W.ggnn_weights(L, D, seed=3, random_bias=True)), False),
        "rgat": (lambda h, adj, c: R.sparse_rgat_layer(h, adj, D, num_heads=2, weights=W.rgat_weights(L, D, D, seed=4)), False),
        "film": (lambda h, adj, c: R.sparse_gnn_film_layer(h, adj, c, D, normalize_by_num_incoming=True, weights=W.film_weights(L, D, D, seed=5, random_ln=True)), True),
        "edge_mlp": (lambda h, adj, c: R.sparse_gnn_edge_mlp_layer(h, adj, c, D, message_aggregation_function="mean",
                                                                  weights=W.edge_mlp_weights(L, D, D, 1, True, seed=6, random_ln=True)), True),
        "rgin": (lambda h, adj, c: R.sparse_rgin_layer(h, adj, D, num_aggr_MLP_hidden_layers=0, message_aggregation_function="sqrt_n",
                                                      weights=W.rgin_weights(L, D, D, 1, 0, False, seed=7, random_ln=True)), False),
        "rgdcn": (lambda h, adj, c: R.sparse_rgdcn_layer(h, adj, c, 2, 4, weights=W.rgdcn_weights(L, 2, 4, seed=8, stddev=0.3)), True),
    }, adj, indeg, h


NAMES = ["rgcn", "rgcn_max_both", "ggnn", "rgat", "film", "edge_mlp", "rgin", "rgdcn"]


def in_degrees(adj, n):
    return np.stack([np.bincount(a[:, 1], minlength=n) for a in adj]).astype(np.float32)


@pytest.mark.parametrize("name", NAMES)
def test_edge_order_within_a_type_does_not_matter(name):
    fns, adj, indeg, h = cases()
    fn, _ = fns[name]
    rng = np.random.default_rng(1)
    shuffled = [a[rng.permutation(a.shape[0])] for a in adj]
    assert R.max_norm_rel_err(fn(h, shuffled, indeg), fn(h, adj, indeg)) < 1e-12


@pytest.mark.parametrize("name", NAMES)
def test_node_relabelling_is_an_equivariance(name):
    fns, adj, indeg, h = cases()
    fn, _ = fns[name]
    perm = np.random.default_rng(2).permutation(V)            # new id of old node i is perm[i]
    h2 = np.empty_like(h); h2[perm] = h
    adj2 = [perm[a].astype(np.int32).reshape(-1, 2) for a in adj]
    out2 = fn(h2, adj2, in_degrees(adj2, V))
    assert R.max_norm_rel_err(out2[perm], fn(h, adj, indeg)) < 1e-12


@pytest.mark.parametrize("name", NAMES)
def test_an_extra_isolated_node_changes_nothing_for_the_others(name):
    fns, adj, indeg, h = cases()
    fn, _ = fns[name]
    h3 = np.concatenate([h, np.full((1, D), 0.37)])
    out3 = fn(h3, adj, in_degrees(adj, V + 1))
    assert R.max_norm_rel_err(out3[:V], fn(h, adj, indeg)) < 1e-12
