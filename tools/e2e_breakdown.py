"""Where the end-to-end step goes: each phase timed with a synchronize on both sides (so no overlap)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W

dev = torch.device("cuda", 0)
G.set_weight_cache(True)
b = batching.ppi_like_batch()
h0 = np.tanh(np.random.default_rng(1).standard_normal((b.num_nodes, 256))).astype(np.float32)
ws = [W.to_torch(W.rgcn_weights(3, 256, 256, seed=2 + 10 * i), dev) for i in range(3)]
pin = lambda a: torch.as_tensor(a).pin_memory()
h_host, adj_host, cnt_host = pin(h0), [pin(np.ascontiguousarray(a)) for a in b.adjacency_lists], pin(b.type_to_num_incoming_edges)
out_host = torch.empty((b.num_nodes, 256)).pin_memory()


def phase(fn, n=50):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6, r

t_h2d, (hd, ad, cd) = phase(lambda: (h_host.to(dev, non_blocking=True), [a.to(dev, non_blocking=True) for a in adj_host], cnt_host.to(dev, non_blocking=True)))
t_plan, p = phase(lambda: G.GraphPlan(ad, b.num_nodes, device=dev))
def layers():
    cur = hd
    for w in ws:
        cur = G.sparse_rgcn_layer(cur, p, cd, 256, activation_function="ReLU", weights=w)
    return cur
t_layers, cur = phase(layers)
t_d2h, _ = phase(lambda: out_host.copy_(cur, non_blocking=True))
def enqueue_only():
    t0 = time.perf_counter(); layers(); return time.perf_counter() - t0
torch.cuda.synchronize()
enq = sorted(enqueue_only() for _ in range(50))[25] * 1e6
print("H2D %.0f us | plan build %.0f us | 3 layers %.0f us (host enqueue alone %.0f us) | D2H %.0f us" % (t_h2d, t_plan, t_layers, enq, t_d2h))
