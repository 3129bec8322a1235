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
