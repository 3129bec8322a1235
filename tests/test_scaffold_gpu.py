"""GPU: the torch scaffold around the engine (RGCN_Model on the PPI task): README parameter count, whole-model
forward parity against the numpy oracle, and a few training steps through the engine's backward."""
import numpy as np
import pytest

from oracle import ref_model
from tf_gnn_samples_b200 import GraphPlan, batching
from tf_gnn_samples_b200.scaffold import RGCNPPIModel

from helpers import assert_parity

pytestmark = pytest.mark.gpu


def model_weights_numpy(model):
    L = model.num_edge_types
    return {
        "projection": model.projection.detach().cpu().numpy(),
        "layers": [{"edge_weights": [w.detach().cpu().numpy() for w in model.layer_weights(l)["edge_weights"]]}
                   for l in range(model.params["graph_num_layers"])],
        "inter_dense": {int(k): v.detach().cpu().numpy() for k, v in model.inter_dense.items()},
        "out_kernel": model.out_kernel.detach().cpu().numpy(), "out_bias": model.out_bias.detach().cpu().numpy(),
    }


def test_parameter_count_matches_readme(cuda_device):
    """README.md:29: 'Model has 699257 parameters' for RGCN on PPI with hidden 256 and 3 layers."""
    assert RGCNPPIModel(device=cuda_device).num_parameters() == 699257


def test_whole_model_forward_parity(cuda_device):
    import torch
    b = batching.ppi_like_batch(num_nodes=600, num_links=9000, seed=12)
    model = RGCNPPIModel(device=cuda_device).eval()
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device)
    with torch.no_grad():
        logits = model(torch.as_tensor(b.node_features).to(cuda_device), plan,
                       torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)).cpu().numpy()
    want = ref_model.rgcn_ppi_logits(b.node_features, b.adjacency_lists, b.type_to_num_incoming_edges, model.params,
                                     model_weights_numpy(model))
    assert_parity(logits, want, "RGCN/PPI whole-model logits", tol=2e-4)   # torch fp32 matmuls around 3 engine layers


def test_training_steps_reduce_the_loss(cuda_device):
    import torch
    torch.manual_seed(0)
    b = batching.ppi_like_batch(num_nodes=500, num_links=7000, seed=13)
    model = RGCNPPIModel(device=cuda_device, params={"learning_rate": 0.003})
    feats = torch.as_tensor(b.node_features).to(cuda_device)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(cuda_device)
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=cuda_device)
    labels = (torch.rand((b.num_nodes, 121), device=cuda_device) < 0.3).float()
    opt = model.make_optimizer()
    losses = [model.train_step(opt, feats, plan, cnt, labels)["loss"] for _ in range(12)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], losses
    for p in model.edge_weights:                                   # the engine's backward reached every GNN kernel
        assert p.grad is not None and float(p.grad.abs().max()) > 0
