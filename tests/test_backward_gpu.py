"""GPU: gradients of the RGCN layer (rgnn_rgcn_backward through torch.autograd) against the analytic float64
oracle gradients (oracle/ref_grads.py, pinned by finite differences in tests/test_oracle_grads.py)."""
import numpy as np
import pytest

from oracle import ref_grads as RG
from oracle import ref_layers as R
from tf_gnn_samples_b200 import GraphPlan, batching, sparse_rgcn_layer, weights as W

from helpers import assert_parity, node_states, tiny_graph

pytestmark = pytest.mark.gpu


def engine_grads(cuda_device, h, adj, indeg, w, g, act, agg, normalize, d_out, T=1):
    import torch
    ht = torch.as_tensor(h).to(cuda_device).requires_grad_(True)
    wt = [torch.as_tensor(k).to(cuda_device).requires_grad_(True) for k in w["edge_weights"]]
    cnt = torch.as_tensor(indeg).to(cuda_device)
    out = sparse_rgcn_layer(ht, adj, cnt, d_out, num_timesteps=T, activation_function=act,
                            message_aggregation_function=agg, normalize_by_num_incoming=normalize,
                            weights={"edge_weights": wt})
    (out * torch.as_tensor(g).to(cuda_device)).sum().backward()
    return out.detach().cpu().numpy(), ht.grad.cpu().numpy(), [k.grad.cpu().numpy() for k in wt]


@pytest.mark.parametrize("act,agg,normalize", [("tanh", "sum", True), ("ReLU", "mean", False), ("elu", "sqrt_n", True),
                                               ("gelu", "sum", True), ("linear", "sum", False), ("selu", "mean", True),
                                               ("leaky_relu", "sum", True)])
def test_rgcn_grads_small(cuda_device, act, agg, normalize):
    adj, indeg = tiny_graph(61, (170, 61, 0, 95), seed=51)
    d_in, d_out = 64, 96
    h = node_states(61, d_in, seed=52) + 0.02
    w = W.rgcn_weights(4, d_in, d_out, seed=53)
    g = np.random.default_rng(54).standard_normal((61, d_out)).astype(np.float32)
    out, d_h, d_ws = engine_grads(cuda_device, h, adj, indeg, w, g, act, agg, normalize, d_out)
    want_h, want_ws = RG.rgcn_layer_grads(h, adj, indeg, g, act, agg, normalize, weights=w)
    assert_parity(out, R.sparse_rgcn_layer(h, adj, indeg, d_out, activation_function=act, message_aggregation_function=agg,
                                           normalize_by_num_incoming=normalize, weights=w), "forward %s" % act)
    assert_parity(d_h, want_h, "d_h %s %s" % (act, agg))
    for l in range(4):
        if adj[l].shape[0] == 0:
            assert np.all(d_ws[l] == 0)
        else:
            assert_parity(d_ws[l], want_ws[l], "d_W[%d] %s %s" % (l, act, agg))


def test_rgcn_grads_ppi_shaped_two_timesteps(cuda_device):
    """hidden 256 on a PPI-shaped graph, two chained timesteps (autograd composes the per-step backward)."""
    b = batching.ppi_like_batch(num_nodes=900, num_links=20000, seed=6)
    D = 256
    h = node_states(b.num_nodes, D, seed=7)
    w = W.rgcn_weights(3, D, D, seed=8)
    g = np.random.default_rng(9).standard_normal((b.num_nodes, D)).astype(np.float32)
    out, d_h, d_ws = engine_grads(cuda_device, h, b.adjacency_lists, b.type_to_num_incoming_edges, w, g, "tanh", "sum", True, D, T=2)
    # oracle: chain rule through the two timesteps
    h1 = R.sparse_rgcn_layer(h, b.adjacency_lists, b.type_to_num_incoming_edges, D, activation_function="tanh", weights=w)
    g1, dw2 = RG.rgcn_layer_grads(h1, b.adjacency_lists, b.type_to_num_incoming_edges, g, "tanh", "sum", True, weights=w)
    g0, dw1 = RG.rgcn_layer_grads(h, b.adjacency_lists, b.type_to_num_incoming_edges, g1, "tanh", "sum", True, weights=w)
    assert_parity(d_h, g0, "d_h two timesteps")
    for l in range(3):
        assert_parity(d_ws[l], dw1[l] + dw2[l], "d_W[%d] two timesteps (shared kernels)" % l)


def test_rgcn_max_aggregation_under_autograd(cuda_device):
    """max aggregation is outside the fused backward kernels: it takes the composed path (gnns/_train.py) and still
    returns gradients (compared with the oracle in test_train_layers_gpu.py)."""
    import torch
    adj, indeg = tiny_graph()
    h = torch.as_tensor(node_states(37, 64)).to(cuda_device).requires_grad_(True)
    w = {"edge_weights": [torch.as_tensor(k).to(cuda_device) for k in W.rgcn_weights(4, 64, 64)["edge_weights"]]}
    out = sparse_rgcn_layer(h, adj, indeg, 64, message_aggregation_function="max", weights=w)
    out.sum().backward()
    assert h.grad is not None and torch.isfinite(h.grad).all()
    with torch.no_grad():   # inference with max aggregation: the fused kernel
        ref = sparse_rgcn_layer(h, adj, indeg, 64, message_aggregation_function="max", weights=w)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_rgcn_grads_with_heavy_segments(cuda_device):
    """Zipf-skewed targets: heavy targets in the forward, heavy (source, type) pairs in the reverse index."""
    b = batching.ppi_like_batch(num_nodes=800, num_links=40000, zipf_targets=True, seed=33)
    D = 64
    h = node_states(b.num_nodes, D, seed=34)
    w = W.rgcn_weights(3, D, D, seed=35)
    g = np.random.default_rng(36).standard_normal((b.num_nodes, D)).astype(np.float32)
    out, d_h, d_ws = engine_grads(cuda_device, h, b.adjacency_lists, b.type_to_num_incoming_edges, w, g, "tanh", "sum", True, D)
    want_h, want_ws = RG.rgcn_layer_grads(h, b.adjacency_lists, b.type_to_num_incoming_edges, g, "tanh", "sum", True, weights=w)
    assert_parity(d_h, want_h, "d_h heavy")
    for l in range(3):
        assert_parity(d_ws[l], want_ws[l], "d_W[%d] heavy" % l)
