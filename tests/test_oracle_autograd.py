"""CPU: pin oracle/ref_autograd.py (torch float64, differentiable) against oracle/ref_layers.py (numpy float64) on the
forward pass of every layer family, and its gradients with torch.autograd.gradcheck (finite differences)."""
import numpy as np
import pytest

from oracle import ref_autograd as A
from oracle import ref_layers as R
from tf_gnn_samples_b200 import weights as W

from helpers import node_states, tiny_graph

torch = pytest.importorskip("torch")

V, L, D = 37, 4, 16


def case(name, T=1):
    adj, indeg = tiny_graph(V, (60, 0, 45, 11), seed=3)
    h = node_states(V, D, seed=4).astype(np.float64)
    if name == "rgcn":
        w = W.rgcn_weights(L, D, D, seed=5)
        np_out = R.sparse_rgcn_layer(h, adj, indeg, D, num_timesteps=T, activation_function="tanh",
                                     message_aggregation_function="mean", weights=w)
        fn = lambda hh, ww: A.sparse_rgcn_layer(hh, adj, torch.as_tensor(indeg, dtype=torch.float64), num_timesteps=T,
                                                activation_function="tanh", message_aggregation_function="mean", weights=ww)
    elif name == "rgcn_both_max":
        w = W.rgcn_weights(L, D, D, seed=5, use_both_source_and_target=True)
        np_out = R.sparse_rgcn_layer(h, adj, indeg, D, activation_function="elu", message_aggregation_function="max",
                                     use_both_source_and_target=True, weights=w)
        fn = lambda hh, ww: A.sparse_rgcn_layer(hh, adj, torch.as_tensor(indeg, dtype=torch.float64), activation_function="elu",
                                                message_aggregation_function="max", use_both_source_and_target=True, weights=ww)
    elif name in ("ggnn_gru", "ggnn_rnn"):
        cell = name.split("_")[1]
        w = W.ggnn_weights(L, D, seed=6, cell=cell, random_bias=True)
        np_out = R.sparse_ggnn_layer(h, adj, D, num_timesteps=T, gated_unit_type=cell, weights=w)
        fn = lambda hh, ww: A.sparse_ggnn_layer(hh, adj, num_timesteps=T, gated_unit_type=cell, weights=ww)
    elif name == "rgat":
        w = W.rgat_weights(L, D, D, seed=7)
        np_out = R.sparse_rgat_layer(h, adj, D, num_timesteps=T, num_heads=4, weights=w)
        fn = lambda hh, ww: A.sparse_rgat_layer(hh, adj, num_timesteps=T, num_heads=4, weights=ww)
    elif name == "film":
        w = W.film_weights(L, D, D, seed=8, num_timesteps=T, random_ln=True)
        np_out = R.sparse_gnn_film_layer(h, adj, indeg, D, num_timesteps=T, activation_function="gelu",
                                         normalize_by_num_incoming=True, weights=w)
        fn = lambda hh, ww: A.sparse_gnn_film_layer(hh, adj, torch.as_tensor(indeg, dtype=torch.float64), num_timesteps=T,
                                                    activation_function="gelu", normalize_by_num_incoming=True, weights=ww)
    elif name == "edge_mlp":
        w = W.edge_mlp_weights(L, D, D, 1, True, seed=9, num_timesteps=T, random_ln=True)
        np_out = R.sparse_gnn_edge_mlp_layer(h, adj, indeg, D, num_timesteps=T, activation_function="leaky_relu",
                                             message_aggregation_function="sqrt_n", weights=w)
        fn = lambda hh, ww: A.sparse_gnn_edge_mlp_layer(hh, adj, None, num_timesteps=T, activation_function="leaky_relu",
                                                        message_aggregation_function="sqrt_n", weights=ww)
    elif name == "rgin":
        w = W.rgin_weights(L, D, D, 1, 1, False, seed=10, num_timesteps=T, random_ln=True)
        np_out = R.sparse_rgin_layer(h, adj, D, num_timesteps=T, activation_function="selu", num_aggr_MLP_hidden_layers=1, weights=w)
        fn = lambda hh, ww: A.sparse_rgin_layer(hh, adj, num_timesteps=T, activation_function="selu", weights=ww)
    else:
        raise KeyError(name)
    return h, w, np_out, fn


NAMES = ["rgcn", "rgcn_both_max", "ggnn_gru", "ggnn_rnn", "rgat", "film", "edge_mlp", "rgin"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("T", [1, 2])
def test_forward_matches_numpy_oracle(name, T):
    if name == "rgcn_both_max" and T == 2:
        pytest.skip("single timestep case")
    h, w, np_out, fn = case(name, T)
    out = fn(torch.as_tensor(h), A.to_torch64(w, requires_grad=False)).numpy()
    assert R.max_norm_rel_err(out, np_out) < 1e-12


@pytest.mark.parametrize("name", ["rgcn", "ggnn_gru", "rgat", "film", "edge_mlp", "rgin"])
def test_gradients_pass_gradcheck(name):
    """Finite differences on a few scalar projections of the output (tanh / smooth activations only: the cases above
    use smooth activations or are evaluated away from kinks by gradcheck's tolerance)."""
    h, w, _, fn = case(name, 1)
    wt = A.to_torch64(w)
    flat = A.flatten(wt)
    ht = torch.as_tensor(h).clone().requires_grad_(True)
    proj = torch.as_tensor(np.random.default_rng(0).standard_normal((V, D)))
    names = list(flat)
    leaves = [flat[n] for n in names]
    assert torch.autograd.gradcheck(lambda hh, *ps: (fn(hh, _rebuild(wt, names, ps)) * proj).sum(), [ht] + leaves,
                                    eps=1e-6, atol=1e-6, rtol=1e-5, nondet_tol=0.0)


def _rebuild(template, names, tensors):
    """The weight container with its leaves replaced (in flatten() order)."""
    it = iter(tensors)

    def walk(x):
        if isinstance(x, dict):
            return {k: walk(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [walk(v) for v in x]
        if x is None:
            return None
        return next(it)
    return walk(template)
