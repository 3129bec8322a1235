"""Where the end-to-end step goes.  Variants of the host->device->host step, each timed as a whole
(wall clock, synchronised at both ends) plus phase timings with a synchronize after each phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W

dev = torch.device("cuda", 0)
G.set_weight_cache(True)
b = batching.ppi_like_batch()
V, L = b.num_nodes, 3
h0 = np.tanh(np.random.default_rng(1).standard_normal((V, 256))).astype(np.float32)
ws = [W.to_torch(W.rgcn_weights(3, 256, 256, seed=2 + 10 * i), dev) for i in range(3)]
pin = lambda a: torch.as_tensor(a).pin_memory()
h_host, adj_host, cnt_host = pin(h0), [pin(np.ascontiguousarray(a)) for a in b.adjacency_lists], pin(b.type_to_num_incoming_edges)
out_host = torch.empty((V, 256)).pin_memory()
sections = [("h", h0)] + [("adj%d" % i, np.ascontiguousarray(a)) for i, a in enumerate(b.adjacency_lists)] + [("cnt", b.type_to_num_incoming_edges)]
offsets, total = {}, 0
for name, arr in sections:
    offsets[name] = (total, arr.nbytes, arr.dtype, arr.shape); total += (arr.nbytes + 255) // 256 * 256
stage_host = torch.empty(total, dtype=torch.uint8).pin_memory()
for name, arr in sections:
    o, nb, _, _ = offsets[name]; stage_host[o:o + nb] = torch.as_tensor(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
stage_dev = torch.empty(total, dtype=torch.uint8, device=dev)
def dev_view(name):
    o, nb, dt, shape = offsets[name]
    return stage_dev[o:o + nb].view(torch.float32 if dt == np.float32 else torch.int32).view(*shape)


def wall(fn, n=60):
    for _ in range(5): fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    ts.sort(); return ts[len(ts) // 2] * 1e6

def h2d_sep(): return h_host.to(dev, non_blocking=True), [a.to(dev, non_blocking=True) for a in adj_host], cnt_host.to(dev, non_blocking=True)
def h2d_staged():
    stage_dev.copy_(stage_host, non_blocking=True); return dev_view("h"), [dev_view("adj%d" % i) for i in range(L)], dev_view("cnt")
def layers3(hd, p, cd):
    cur = hd
    for w in ws: cur = G.sparse_rgcn_layer(cur, p, cd, 256, activation_function="ReLU", weights=w)
    return cur
def stack(hd, p, cd): return G.rgcn_layer_stack(hd, p, cd, ws, activation_function="ReLU")

def step(h2d, validate, run, check):
    hd, ad, cd = h2d()
    p = G.GraphPlan(ad, V, device=dev, validate=validate)
    cur = run(hd, p, cd)
    out_host.copy_(cur, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    if check: p.check()
    p.close()

print("H2D separate %.0f us | H2D staged %.0f us" % (wall(h2d_sep), wall(h2d_staged)))
hd, ad, cd = h2d_staged(); torch.cuda.synchronize()
print("plan validate=True %.0f us | validate=False %.0f us" % (wall(lambda: G.GraphPlan(ad, V, device=dev).close()), wall(lambda: G.GraphPlan(ad, V, device=dev, validate=False).close())))
p = G.GraphPlan(ad, V, device=dev)
print("3 layer calls %.0f us | stack call %.0f us | D2H %.0f us | check %.0f us" % (wall(lambda: layers3(hd, p, cd)), wall(lambda: stack(hd, p, cd)),
      wall(lambda: out_host.copy_(hd, non_blocking=True)), wall(lambda: p.check())))
for name, args in [("separate+validate+3calls", (h2d_sep, True, layers3, False)), ("staged+validate+3calls", (h2d_staged, True, layers3, False)),
                   ("staged+deferred+3calls", (h2d_staged, False, layers3, True)), ("staged+deferred+stack", (h2d_staged, False, stack, True)),
                   ("staged+deferred+stack no check", (h2d_staged, False, stack, False)), ("separate+deferred+stack", (h2d_sep, False, stack, True))]:
    print("step %-32s %.0f us" % (name, wall(lambda: step(*args))))
