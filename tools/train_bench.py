"""Whole-model throughput of RGCN on a PPI-shaped training batch through the torch scaffold + the engine:
forward-only (the reference's validation pass) and forward+backward+Adam (its training pass).  The reference's
README log reports 1,952,084 edges/s (train) and 3,098,674 edges/s (valid) for this model (README.md:34-35,
hardware per the README tables: V100, whole pipeline incl. Python batching) -- an order-of-magnitude anchor only.
Device-timed with CUDA events, median of N; batch = 5 packed PPI-shaped graphs (~max_nodes_in_batch 12,500)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching
from tf_gnn_samples_b200.scaffold import RGCNPPIModel

dev = torch.device("cuda", 0)
b = batching.ppi_like_batch(num_graphs=5, seed=0)
model = RGCNPPIModel(device=dev)
feats = torch.as_tensor(b.node_features).to(dev)
cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(dev)
labels = (torch.rand((b.num_nodes, 121), device=dev) < 0.3).float()
plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
opt = model.make_optimizer()


def timed(fn, n=40):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def fwd():
    model.eval()
    with torch.no_grad():
        return model(feats, plan, cnt)

ms_train = timed(lambda: model.train_step(opt, feats, plan, cnt, labels))
ms_eval = timed(fwd)
print(json.dumps({"model": "RGCN PPI hidden=256 3 layers (699,257 params)", "V": b.num_nodes, "M": b.num_edges,
                  "train_step_ms": ms_train, "train_edges_per_s": b.num_edges / (ms_train * 1e-3),
                  "eval_forward_ms": ms_eval, "eval_edges_per_s": b.num_edges / (ms_eval * 1e-3),
                  "reference_readme_v100": {"train_edges_per_s": 1952084, "valid_edges_per_s": 3098674}}))

# ---- the other families through the general scaffold: QM9-shaped batch (2,000 molecule graphs), hidden 128, 2 layers ----
from tf_gnn_samples_b200.scaffold import SparseGraphModel
qb = batching.qm9_like_batch(num_graphs=2000, seed=1, add_self_loop_edges=True)
qplan = G.GraphPlan(qb.adjacency_lists, qb.num_nodes, device=dev)
qfeats = torch.as_tensor(qb.node_features).to(dev)
qcnt = torch.as_tensor(qb.type_to_num_incoming_edges).to(dev)
sizes = np.diff(qb.graph_node_offsets)
gl = torch.as_tensor(np.repeat(np.arange(qb.num_graphs), sizes)).to(dev)
tg = torch.randn((1, qb.num_graphs), device=dev)
rows = {}
for kind in ["ggnn", "rgat", "rgin", "gnn-edge-mlp", "gnn-film", "rgcn"]:
    m = SparseGraphModel(kind, "qm9", num_edge_types=5, feature_size=15, device=dev,
                         params={"graph_num_layers": 2, "graph_layer_input_dropout_keep_prob": 1.0,
                                 "graph_num_timesteps_per_layer": 4 if kind == "ggnn" else 1})
    o = m.make_optimizer()
    def fwd_q():
        m.eval()
        with torch.no_grad():
            return m(qfeats, qplan, qcnt, gl, qb.num_graphs)
    t_train = timed(lambda: m.train_step_async(o, qfeats, qplan, qcnt, tg, gl, qb.num_graphs), n=20)
    t_eval = timed(fwd_q, n=20)
    rows[kind] = {"train_step_ms": t_train, "eval_forward_ms": t_eval, "train_edges_per_s": qb.num_edges / (t_train * 1e-3),
                  "eval_edges_per_s": qb.num_edges / (t_eval * 1e-3)}
print(json.dumps({"qm9_shaped": {"graphs": qb.num_graphs, "V": qb.num_nodes, "M": qb.num_edges, "L": 5, "hidden": 128,
                                 "layers": 2, "models": rows}}))
