"""Time one PPI-shaped RGCN layer (cold L2, CUDA events) -- run once per environment setting:
   RGNN_SEG_GROUP=2|4|8  RGNN_SEG_COLS=128|256  RGNN_GEMM_IMPL=mma|tc   python tools/layer_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W, ops

dev = torch.device("cuda", 0)
b = batching.ppi_like_batch()
h = torch.as_tensor(np.tanh(np.random.default_rng(1).standard_normal((b.num_nodes, 256))).astype(np.float32)).to(dev)
cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(dev)
plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
w = W.to_torch(W.rgcn_weights(3, 256, 256), dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=60):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

layer = lambda: G.sparse_rgcn_layer(h, plan, cnt, 256, activation_function="ReLU", weights=w)
graph = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    layer(); layer()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
with torch.cuda.graph(graph):
    out = layer()
msgs = torch.randn(b.num_edges, 256, device=dev)
t_layer = timeit(graph.replay)
t_seg_only = timeit(lambda: ops.segment_aggregate(plan, msgs, "sum"))
print("env GROUP=%s COLS=%s GEMM=%s : layer(graph) %.1f us   segment_aggregate[M,256] %.1f us" % (
    os.environ.get("RGNN_SEG_GROUP", "-"), os.environ.get("RGNN_SEG_COLS", "-"), os.environ.get("RGNN_GEMM_IMPL", "tc"),
    t_layer, t_seg_only))
