"""CPU: reference snapshot compatibility -- tf-variable-name sorting for every layer family and a
save -> load round trip of the RGCN/PPI scaffold in the reference's pickle structure."""
import pickle

import numpy as np
import pytest

from tf_gnn_samples_b200 import checkpoint as C


def fake(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)


def test_sorting_of_every_layer_family():
    w = {
        "graph_model/dense/kernel:0": fake((50, 64), 0),
        # layer 0: RGAT-style
        "graph_model/gnn_layer_0/Edge_0_Weight/kernel:0": fake((64, 64), 1),
        "graph_model/gnn_layer_0/Edge_1_Weight/kernel:0": fake((64, 64), 2),
        "graph_model/gnn_layer_0/Edge_0_Attention_Parameters:0": fake((128,), 3),
        "graph_model/gnn_layer_0/Edge_1_Attention_Parameters:0": fake((128,), 4),
        "graph_model/gnn_layer_0/Dense/kernel:0": fake((64, 64), 5),
        # layer 1: FiLM-style with two timesteps + inter-layer norm
        "graph_model/gnn_layer_1/Edge_0_Weight/kernel:0": fake((64, 64), 6),
        "graph_model/gnn_layer_1/Edge_0_FiLM_Computations/kernel:0": fake((64, 128), 7),
        "graph_model/gnn_layer_1/LayerNorm/gamma:0": fake((64,), 8),
        "graph_model/gnn_layer_1/LayerNorm/beta:0": fake((64,), 9),
        "graph_model/gnn_layer_1/LayerNorm_1/gamma:0": fake((64,), 10),
        "graph_model/gnn_layer_1/LayerNorm_1/beta:0": fake((64,), 11),
        "graph_model/gnn_layer_1/LayerNorm_2/gamma:0": fake((64,), 12),
        "graph_model/gnn_layer_1/LayerNorm_2/beta:0": fake((64,), 13),
        # layer 2: RGIN / Edge-MLP style
        "graph_model/gnn_layer_2/Edge_0_MLP/dense/kernel:0": fake((128, 64), 14),
        "graph_model/gnn_layer_2/Edge_0_MLP/dense_1/kernel:0": fake((64, 64), 15),
        "graph_model/gnn_layer_2/Edge_1_MLP/dense/kernel:0": fake((128, 64), 16),
        "graph_model/gnn_layer_2/Edge_1_MLP/dense_1/kernel:0": fake((64, 64), 17),
        "graph_model/gnn_layer_2/Aggregation_MLP/dense/kernel:0": fake((64, 64), 18),
        # layer 3: GGNN
        "graph_model/gnn_layer_3/Edge_0_Weight/kernel:0": fake((64, 64), 19),
        "graph_model/gnn_layer_3/gru_cell/kernel:0": fake((64, 192), 20),
        "graph_model/gnn_layer_3/gru_cell/recurrent_kernel:0": fake((64, 192), 21),
        "graph_model/gnn_layer_3/gru_cell/bias:0": fake((192,), 22),
        # layer 4: RGDCN
        "graph_model/gnn_layer_4/Edge_0_Channel_0_Weight_Computation/kernel:0": fake((16, 256), 30),
        "graph_model/gnn_layer_4/Edge_0_Channel_1_Weight_Computation/kernel:0": fake((16, 256), 31),
        "graph_model/gnn_layer_4/Edge_1_Channel_0_Weight_Computation/kernel:0": fake((16, 256), 32),
        "graph_model/gnn_layer_4/Edge_1_Channel_1_Weight_Computation/kernel:0": fake((16, 256), 33),
        # task head, optimizer slots, something unknown
        "out_layer_task/dense_1/kernel:0": fake((64, 121), 23),
        "out_layer_task/dense_1/bias:0": fake((121,), 24),
        "graph_model/gnn_layer_0/Edge_0_Weight/kernel/Adam:0": fake((64, 64), 25),
        "graph_model/gnn_layer_0/Edge_0_Weight/kernel/Adam_1:0": fake((64, 64), 26),
        "beta1_power:0": np.float32(0.9),
        "graph_model/gnn_layer_0/Mystery/kernel:0": fake((3, 3), 27),
    }
    s = C.sort_variables(w)
    assert s["layer_indices"] == [0, 1, 2, 3, 4]
    l0, l1, l2, l3, l4 = s["layers"]
    assert [len(c) for c in l4["channel_weights"]] == [2, 2]
    np.testing.assert_array_equal(l4["channel_weights"][1][0], w["graph_model/gnn_layer_4/Edge_1_Channel_0_Weight_Computation/kernel:0"])
    assert len(l0["edge_weights"]) == 2 and len(l0["attention"]) == 2 and l0["inter_dense"].shape == (64, 64)
    np.testing.assert_array_equal(l0["attention"][1], w["graph_model/gnn_layer_0/Edge_1_Attention_Parameters:0"])
    l1 = C.split_layer_norms(l1, num_timesteps=2)
    assert len(l1["ln_gamma"]) == 2 and len(l1["ln_beta"]) == 2
    np.testing.assert_array_equal(l1["inter_ln_gamma"], w["graph_model/gnn_layer_1/LayerNorm_2/gamma:0"])
    np.testing.assert_array_equal(l1["film_weights"][0], w["graph_model/gnn_layer_1/Edge_0_FiLM_Computations/kernel:0"])
    assert [len(m) for m in l2["edge_mlps"]] == [2, 2] and l2["edge_mlps"][1][0].shape == (128, 64)
    np.testing.assert_array_equal(l2["edge_mlps"][0][1], w["graph_model/gnn_layer_2/Edge_0_MLP/dense_1/kernel:0"])
    assert len(l2["aggr_mlp"]) == 1
    assert l3["cell"]["kind"] == "gru" and l3["cell"]["recurrent_kernel"].shape == (64, 192)
    assert s["unused"] == ["graph_model/gnn_layer_0/Mystery/kernel:0"]
    sc = C.scaffold_variables(s["outside"], feature_size=50, hidden_size=64)
    assert sc["projection"].shape == (50, 64)
    assert len(sc["head"]) == 1 and sc["head"][0]["bias"].shape == (121,)
    assert sc["other"] == {}                                        # Adam slots / beta powers were dropped


def test_snapshot_round_trip_through_the_reference_pickle_structure(tmp_path):
    torch = pytest.importorskip("torch")
    from tf_gnn_samples_b200.scaffold import RGCNPPIModel
    a = RGCNPPIModel(device="cpu", params={"random_seed": 1, "graph_inter_layer_norm": True})
    b = RGCNPPIModel(device="cpu", params={"random_seed": 2, "graph_inter_layer_norm": True})
    path = str(tmp_path / "RGCN_PPI_best_model.pickle")
    C.save_reference_checkpoint(path, "RGCN", "PPI", a.params, {"max_nodes_in_batch": 12500}, a.to_reference_weights())
    with open(path, "rb") as f:                                     # the structure of save_model (:98-105)
        raw = pickle.load(f)
    assert set(raw) == {"model_class", "task_class", "model_params", "task_params", "task_metadata", "weights"}
    assert "gnn_layer_2/Edge_1_Weight/kernel:0" in raw["weights"] and "gnn_layer_0/Dense/kernel:0" in raw["weights"]
    ck = C.load_reference_checkpoint(path)
    assert ck.model_class == "RGCN" and ck.model_params["hidden_size"] == 256
    extra = dict(ck.weights)
    extra["gnn_layer_0/Edge_0_Weight/kernel/Adam:0"] = np.zeros((256, 256), np.float32)     # optimizer slots are ignored
    assert b.load_reference_weights(extra) == []
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), n
    assert sum(v.size for v in ck.weights.values()) == a.num_parameters()


def test_missing_or_misshapen_variables_are_errors():
    pytest.importorskip("torch")
    from tf_gnn_samples_b200.scaffold import RGCNPPIModel
    m = RGCNPPIModel(device="cpu")
    w = m.to_reference_weights()
    bad = dict(w)
    del bad["gnn_layer_1/Edge_2_Weight/kernel:0"]
    with pytest.raises(KeyError):
        m.load_reference_weights(bad)
    bad = dict(w)
    bad["gnn_layer_1/Edge_2_Weight/kernel:0"] = np.zeros((256, 128), np.float32)
    with pytest.raises(ValueError):
        m.load_reference_weights(bad)


def test_unknown_classes_inside_a_snapshot_do_not_block_loading():
    class Vocabulary:                                               # stands for dpu_utils' class in VarMisuse metadata
        pass
    Vocabulary.__module__ = "dpu_utils_not_installed.mlutils"
    Vocabulary.__qualname__ = "Vocabulary"
    import sys, types
    mod = types.ModuleType("dpu_utils_not_installed.mlutils")
    mod.Vocabulary = Vocabulary
    sys.modules["dpu_utils_not_installed"] = types.ModuleType("dpu_utils_not_installed")
    sys.modules["dpu_utils_not_installed.mlutils"] = mod
    blob = pickle.dumps({"model_class": "GGNN", "task_class": "VarMisuse", "model_params": {}, "task_params": {},
                         "task_metadata": {"vocab": Vocabulary()}, "weights": {"a:0": np.ones(3, np.float32)}})
    del sys.modules["dpu_utils_not_installed.mlutils"], sys.modules["dpu_utils_not_installed"]
    ck = C.load_reference_checkpoint(blob)
    assert ck.task_class == "VarMisuse" and ck.weights["a:0"].shape == (3,)
    with pytest.raises(ValueError):
        C.load_reference_checkpoint(pickle.dumps({"no": "weights"}))
