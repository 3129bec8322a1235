"""CPU: SURVEY.md Appendix C -- the parity-trap checklist, one test per item.  Each checks the oracle's behaviour (what the
GPU parity tests compare against) and, where the trap is a default or a signature, the product's layer functions too."""
import inspect

import numpy as np
import pytest

import tf_gnn_samples_b200 as G
from oracle import ref_layers as R
from tf_gnn_samples_b200 import weights as W
from tf_gnn_samples_b200.scaffold import model_default_params, rgcn_ppi_default_params

from helpers import node_states, tiny_graph

V, D, L = 23, 8, 3


def setup():
    adj, indeg = tiny_graph(V, (40, V, 25), seed=61)
    return adj, indeg, node_states(V, D, seed=62).astype(np.float64)


def default(fn, name):
    return inspect.signature(fn).parameters[name].default


def test_01_rgcn_normalisation_on_by_default_with_per_type_in_degree():
    adj, indeg, h = setup()
    w = W.rgcn_weights(L, D, D, seed=1)
    assert default(G.sparse_rgcn_layer, "normalize_by_num_incoming") is True and default(R.sparse_rgcn_layer, "normalize_by_num_incoming") is True
    out = R.sparse_rgcn_layer(h, adj, indeg, D, activation_function=None, weights=w)
    want = np.zeros((V, D))
    for l, a in enumerate(adj):                                              # scale uses the PER-TYPE in-degree of the target
        for (u, v) in a:
            want[v] += (h[u] @ w["edge_weights"][l].astype(np.float64)) / (float(indeg[l][v]) + 1e-7)
    np.testing.assert_allclose(out, want, rtol=1e-10, atol=1e-12)
    assert model_default_params("rgcn").get("normalize_by_num_incoming") is None   # the adapter never overrides it (rgcn_model.py:36-44)


def test_02_film_and_edge_mlp_normalisation_off_by_default():
    for fn in (G.sparse_gnn_film_layer, R.sparse_gnn_film_layer, G.sparse_gnn_edge_mlp_layer, R.sparse_gnn_edge_mlp_layer):
        assert default(fn, "normalize_by_num_incoming") is False
    assert model_default_params("gnn-film")["normalize_messages_by_num_incoming"] is False
    assert "normalize_by_num_incoming" not in model_default_params("gnn-edge-mlp")          # never plumbed (gnn_edge_mlp_model.py:38-48)


def test_03_activation_placement():
    """FiLM / Edge-MLP / RGIN apply the activation per message BEFORE the sum; RGCN / RGAT after; GGNN none on messages."""
    adj, indeg, h = setup()
    w = W.film_weights(L, D, D, seed=2)
    for l in range(L):                                                       # gamma = 1, beta = 0 through a constant input feature
        f = np.zeros((D, 2 * D), np.float32); f[0, :D] = 1.0
        w["film_weights"][l] = f
    h1 = h.copy(); h1[:, 0] = 1.0
    film = R.sparse_gnn_film_layer(h1, adj, indeg, D, activation_function="relu", weights=w)
    tgt = np.concatenate([a[:, 1] for a in adj])
    msgs = np.concatenate([h1[a[:, 0]] @ w["edge_weights"][l].astype(np.float64) for l, a in enumerate(adj)])
    inside = R.layer_norm(R.unsorted_segment_sum(np.maximum(msgs, 0), tgt, V), np.ones(D), np.zeros(D))
    outside = R.layer_norm(np.maximum(R.unsorted_segment_sum(msgs, tgt, V), 0), np.ones(D), np.zeros(D))
    np.testing.assert_allclose(film, inside, rtol=1e-9, atol=1e-9)
    assert np.abs(film - outside).max() > 1e-3
    rw = W.rgcn_weights(L, D, D, seed=3)
    rgcn = R.sparse_rgcn_layer(h, adj, indeg, D, activation_function="relu", normalize_by_num_incoming=False, weights=rw)
    m = np.concatenate([h[a[:, 0]] @ rw["edge_weights"][l].astype(np.float64) for l, a in enumerate(adj)])
    np.testing.assert_allclose(rgcn, np.maximum(R.unsorted_segment_sum(m, tgt, V), 0), rtol=1e-10, atol=1e-12)   # after the sum


def test_04_edge_mlp_hidden_activation_is_elu_rgin_uses_activation_fn():
    adj, indeg, h = setup()
    w = W.edge_mlp_weights(L, D, D, 1, True, seed=4)
    a_relu = R.sparse_gnn_edge_mlp_layer(h, adj, indeg, D, activation_function="relu", weights=w)
    # recompute by hand with ELU inside the MLP and ReLU on the message
    elu = R.get_activation("elu")
    tgt = np.concatenate([a[:, 1] for a in adj])
    msgs = np.concatenate([elu(np.concatenate([h[a[:, 0]], h[a[:, 1]]], 1) @ w["edge_mlps"][l][0].astype(np.float64)) @ w["edge_mlps"][l][1].astype(np.float64)
                           for l, a in enumerate(adj)])
    want = R.layer_norm(R.unsorted_segment_sum(np.maximum(msgs, 0), tgt, V), np.ones(D), np.zeros(D))
    np.testing.assert_allclose(a_relu, want, rtol=1e-9, atol=1e-9)
    rw = W.rgin_weights(L, D, D, 1, None, False, seed=5)
    tanh_out = R.sparse_rgin_layer(h, adj, D, activation_function="tanh", weights=rw)
    msgs = np.concatenate([np.tanh(np.tanh(h[a[:, 0]] @ rw["edge_mlps"][l][0].astype(np.float64)) @ rw["edge_mlps"][l][1].astype(np.float64))
                           for l, a in enumerate(adj)])
    want = R.layer_norm(np.tanh(R.unsorted_segment_sum(msgs, tgt, V)), np.ones(D), np.zeros(D))
    np.testing.assert_allclose(tanh_out, want, rtol=1e-9, atol=1e-9)


def test_05_gru_hard_sigmoid_and_roles():
    """inputs = aggregated messages, state = h (ggnn.py:92); recurrent activation hard_sigmoid = clip(0.2 x + 0.5, 0, 1)."""
    np.testing.assert_allclose(R.hard_sigmoid(np.array([-3.0, -2.5, 0.0, 1.0, 2.5, 9.0])), [0, 0, 0.5, 0.7, 1, 1])
    adj, indeg, h = setup()
    w = W.ggnn_weights(L, D, seed=6, random_bias=True)
    out = R.sparse_ggnn_layer(h, adj, D, weights=w)
    tgt = np.concatenate([a[:, 1] for a in adj])
    m = R.unsorted_segment_sum(np.concatenate([h[a[:, 0]] @ w["edge_weights"][l].astype(np.float64) for l, a in enumerate(adj)]), tgt, V)
    c = {k: v.astype(np.float64) for k, v in w["cell"].items()}
    z = R.hard_sigmoid(m @ c["kernel"][:, :D] + c["bias"][:D] + h @ c["recurrent_kernel"][:, :D])
    r = R.hard_sigmoid(m @ c["kernel"][:, D:2 * D] + c["bias"][D:2 * D] + h @ c["recurrent_kernel"][:, D:2 * D])
    hh = np.tanh(m @ c["kernel"][:, 2 * D:] + c["bias"][2 * D:] + (r * h) @ c["recurrent_kernel"][:, 2 * D:])
    np.testing.assert_allclose(out, z * h + (1 - z) * hh, rtol=1e-10, atol=1e-12)


def test_06_layer_norm_eps_and_biased_variance():
    x = np.array([[1.0, 1.0, 1.0, 1.0], [0.0, 2.0, 4.0, 6.0]])
    out = R.layer_norm(x, np.ones(4), np.zeros(4))
    assert np.all(out[0] == 0.0)                                             # eps 1e-12: a constant row is not blown up
    np.testing.assert_allclose(out[1], (x[1] - 3.0) / np.sqrt(5.0 + 1e-12))  # biased variance = 5


def test_07_leaky_relu_slope_in_rgat_logits():
    f = R.get_activation("leaky_relu")
    np.testing.assert_allclose(f(np.array([-2.0, 3.0])), [-0.4, 3.0])


def test_08_rgat_softmax_spans_all_edge_types_and_empty_targets_are_zero():
    adj, indeg = tiny_graph(V, (40, 0, 25), seed=63, duplicates=False)        # last 3 nodes never targets
    h = node_states(V, D, seed=64).astype(np.float64)
    w = W.rgat_weights(3, D, D, seed=7)
    out = R.sparse_rgat_layer(h, adj, D, num_heads=2, activation_function=None, weights=w)
    assert np.all(out[-3:] == 0.0)
    # attention of one target over the union of its incoming edges of ALL types sums to one
    tgt = np.concatenate([a[:, 1] for a in adj])
    logits = np.random.default_rng(0).standard_normal(tgt.shape[0])
    att = np.exp(R.unsorted_segment_log_softmax(logits, tgt, V))
    s = R.unsorted_segment_sum(att, tgt, V)
    np.testing.assert_allclose(s[np.bincount(tgt, minlength=V) > 0], 1.0, rtol=1e-12)


def test_09_in_degrees_arrive_as_float32():
    from tf_gnn_samples_b200 import batching
    b = batching.ppi_like_batch(num_nodes=50, num_links=200)
    assert b.type_to_num_incoming_edges.dtype == np.float32                  # tasks/sparse_graph_task.py:145
    assert np.float32(1.0) / (np.float32(3.0) + np.float32(1e-7)) == np.float32(1.0 / 3.0)   # 1e-7 vanishes next to c >= 1 in fp32


def test_10_adjacency_columns_are_source_then_target():
    h = np.array([[1.0, 0.0], [0.0, 1.0], [5.0, 5.0]])
    adj = [np.array([[0, 2]], np.int32)]                                     # message flows 0 -> 2
    out = R.sparse_rgcn_layer(h, adj, np.array([[0, 0, 1.0]]), 2, activation_function=None, normalize_by_num_incoming=False,
                              weights={"edge_weights": [np.eye(2)]})
    np.testing.assert_allclose(out, [[0, 0], [0, 0], [1, 0]])


def test_11_scaffold_dense_after_layer_zero():
    """models/sparse_graph_model.py:194: layer_idx % graph_dense_between_every_num_gnn_layers == 0 fires for layer 0 even at 10000."""
    p = rgcn_ppi_default_params()
    assert p["graph_dense_between_every_num_gnn_layers"] == 10000
    from tf_gnn_samples_b200.scaffold import RGCNPPIModel
    m = RGCNPPIModel(device="cpu")
    assert list(m.inter_dense.keys()) == ["0"] and m.num_parameters() == 699257


def test_12_benchmark_config_is_the_readme_run_not_the_json_hypers():
    p = rgcn_ppi_default_params()
    assert (p["hidden_size"], p["graph_num_layers"]) == (256, 3)             # README.md:29-35, not PPI_RGCN.json (320 / 4)


def test_13_error_metric_is_max_norm_relative():
    a = np.array([1e-9, 1.0]); b = np.array([2e-9, 1.0])
    assert R.max_norm_rel_err(a, b) == pytest.approx(1e-9)                   # element-wise relative would say 0.5
