"""Where the RGCN-PPI training step spends its device time: CUDA-event brackets around the phases of
RGCNPPIModel.train_step plus the engine's backward call alone (5 packed PPI-shaped graphs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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


def med(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
model.train()
state = {}

def fwd():
    state["logits"] = model(feats, plan, cnt)
def loss():
    state["m"] = model.task_metrics(state["logits"], labels)
def bwd():
    opt.zero_grad(set_to_none=True)
    state["m"]["loss"].backward(retain_graph=True)

fwd(); loss()
res["forward_train_mode_ms"] = med(fwd)
res["loss_metrics_ms"] = med(loss)
res["backward_ms"] = med(bwd)
res["optimizer_step_ms"] = med(lambda: opt.step())
res["train_step_ms"] = med(lambda: model.train_step(opt, feats, plan, cnt, labels))

# the engine's layer backward alone (one layer, hidden 256)
h = torch.randn(b.num_nodes, 256, device=dev, requires_grad=True)
ws = [torch.nn.Parameter(torch.randn(256, 256, device=dev) * 0.05) for _ in range(3)]
out = G.sparse_rgcn_layer(h, plan, cnt, 256, activation_function="ReLU", weights={"edge_weights": ws})
g = torch.randn_like(out)
res["layer_forward_ms"] = med(lambda: G.sparse_rgcn_layer(h, plan, cnt, 256, activation_function="ReLU", weights={"edge_weights": ws}))
res["layer_backward_ms"] = med(lambda: torch.autograd.grad(out, [h] + ws, g, retain_graph=True))
res["layer_backward_h_only_ms"] = med(lambda: torch.autograd.grad(out, [h], g, retain_graph=True))
print(json.dumps(res))
