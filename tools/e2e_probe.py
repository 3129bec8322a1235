import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W
dev = torch.device("cuda", 0)
G.set_weight_cache(True)
b = batching.ppi_like_batch(); V, L = b.num_nodes, 3
h0 = np.tanh(np.random.default_rng(1).standard_normal((V, 256))).astype(np.float32)
ws = [W.to_torch(W.rgcn_weights(3, 256, 256, seed=2 + 10 * i), dev) for i in range(3)]
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
sync = torch.cuda.synchronize
def phases(clone):
    acc = np.zeros(6)
    for it in range(40):
        sync(); t = [time.perf_counter()]
        stage_dev.copy_(stage_host, non_blocking=True); sync(); t.append(time.perf_counter())
        hd, cd, ad = dev_view("h"), dev_view("cnt"), [dev_view("adj%d" % i) for i in range(L)]
        if clone: hd, cd, ad = hd.clone(), cd.clone(), [a.clone() for a in ad]
        sync(); t.append(time.perf_counter())
        p = G.GraphPlan(ad, V, device=dev, validate=False); sync(); t.append(time.perf_counter())
        cur = G.rgcn_layer_stack(hd, p, cd, ws, activation_function="ReLU"); sync(); t.append(time.perf_counter())
        out_host.copy_(cur, non_blocking=True); sync(); t.append(time.perf_counter())
        p.check(); p.close(); sync(); t.append(time.perf_counter())
        if it >= 5: acc += np.diff(t)
    return acc / 35 * 1e6
for clone in (False, True):
    a = phases(clone)
    print("clone=%s: H2D %.0f | views/clone %.0f | plan %.0f | stack %.0f | D2H %.0f | check+close %.0f | sum %.0f us" % ((clone,) + tuple(a) + (a.sum(),)))
