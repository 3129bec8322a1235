"""GPU: the general scaffold (SparseGraphModel) -- every layer family on 200 REAL QM9 validation molecules
(tests/golden/qm9_valid_subset.json.gz, loaded through batching.qm9_batch like tasks/qm9_task.py) against the numpy
whole-model oracle, the QM9 gated-sum regression head, and training steps through each family's differentiable path."""
import os

import numpy as np
import pytest

from oracle import ref_model
from tf_gnn_samples_b200 import GraphPlan, batching
from tf_gnn_samples_b200.scaffold import SparseGraphModel

from helpers import assert_parity, assert_parity_8c

pytestmark = pytest.mark.gpu
KINDS = ["rgcn", "ggnn", "rgat", "rgin", "gnn-edge-mlp", "gnn-film"]


def qm9_inputs(device, **kw):
    import torch
    recs = batching.load_qm9_jsonl(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qm9_valid_subset.json.gz"))
    b, gl, tg = batching.qm9_batch(recs, **kw)
    plan = GraphPlan(b.adjacency_lists, b.num_nodes, device=device)
    return (b, gl, tg, plan, torch.as_tensor(b.node_features).to(device), torch.as_tensor(b.type_to_num_incoming_edges).to(device),
            torch.as_tensor(gl).to(device), torch.as_tensor(tg).to(device))


def to_numpy(obj):
    if isinstance(obj, dict):
        return {k: to_numpy(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_numpy(v) for v in obj]
    if obj is None or isinstance(obj, str):
        return obj
    return obj.detach().cpu().numpy()


@pytest.mark.parametrize("kind", KINDS + ["rgdcn"])
def test_whole_model_on_real_qm9_molecules(cuda_device, kind):
    import torch
    b, gl, tg, plan, feats, cnt, gl_d, tg_d = qm9_inputs(cuda_device)
    params = {"graph_num_layers": 3, "hidden_size": 64, "graph_num_timesteps_per_layer": 2 if kind == "ggnn" else 1,
              "graph_layer_input_dropout_keep_prob": 1.0, "graph_residual_connection_every_num_layers": 2,
              "graph_dense_between_every_num_gnn_layers": 2, "random_seed": 3}
    model = SparseGraphModel(kind, "qm9", num_edge_types=5, feature_size=15, params=params, task_ids=(0, 4), device=cuda_device).eval()
    with torch.no_grad():
        final = model.node_representations(feats, plan, cnt)
        out = model(feats, plan, cnt, gl_d, b.num_graphs)
    want_final = ref_model.node_representations(model.kind, b.node_features, b.adjacency_lists, b.type_to_num_incoming_edges,
                                                model.params, to_numpy(model.projection), to_numpy(model.layers))
    want_final32 = ref_model.node_representations(model.kind, b.node_features, b.adjacency_lists, b.type_to_num_incoming_edges,
                                                  model.params, to_numpy(model.projection), to_numpy(model.layers), dtype=np.float32)
    assert_parity_8c(final.cpu().numpy(), want_final, want_final32, "%s node representations on QM9" % kind)
    want = ref_model.qm9_outputs(want_final, b.node_features, gl, b.num_graphs, to_numpy(model.head))
    assert out.shape == (2, b.num_graphs)
    assert_parity(out.cpu().numpy(), want, "%s QM9 per-graph outputs" % kind, tol=1e-4)
    tg2 = np.stack([tg[0], -tg[0]])
    m = model.task_metrics(out, torch.as_tensor(tg2).to(cuda_device))
    ref_m = ref_model.qm9_metrics(want, tg2, (0, 4))
    for k in ref_m:
        assert abs(float(m[k]) - ref_m[k]) <= 5e-4 * max(1.0, abs(ref_m[k])), (k, float(m[k]), ref_m[k])


@pytest.mark.parametrize("kind", KINDS)
def test_training_steps_on_qm9(cuda_device, kind):
    """A few Adam steps through each family's differentiable path reduce the regression loss; every parameter that
    the forward uses receives a finite gradient."""
    import torch
    b, gl, tg, plan, feats, cnt, gl_d, tg_d = qm9_inputs(cuda_device)
    params = {"graph_num_layers": 2, "hidden_size": 64, "graph_layer_input_dropout_keep_prob": 1.0, "learning_rate": 0.003,
              "graph_rnn_cell": "GRU"}
    model = SparseGraphModel(kind, "qm9", num_edge_types=5, feature_size=15, params=params, device=cuda_device)
    opt = model.make_optimizer()
    losses = [model.train_step(opt, feats, plan, cnt, tg_d, gl_d, b.num_graphs)["loss"] for _ in range(15)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], (kind, losses)
    missing = [n for n, q in model.named_parameters() if q.grad is None]
    assert not missing, missing
    assert all(bool(torch.isfinite(q.grad).all()) for q in model.parameters())


def test_ppi_head_and_reference_defaults(cuda_device):
    """RGCN under the general scaffold has the README's parameter count; the adapters' default_params are the reference's."""
    from tf_gnn_samples_b200.scaffold import model_default_params
    m = SparseGraphModel("rgcn", "ppi", 3, 50, params={"graph_num_layers": 3, "hidden_size": 256}, device=cuda_device)
    assert m.num_parameters() == 699257
    assert model_default_params("GGNN")["graph_rnn_cell"] == "GRU" and model_default_params("RGAT")["num_heads"] == 4
    assert model_default_params("GNN-Edge-MLP")["graph_activation_function"] == "gelu"
    assert model_default_params("rgin")["graph_inter_layer_norm"] is True
    with pytest.raises(ValueError):
        model_default_params("transformer")
