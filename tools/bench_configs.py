"""Device-timed layer rates for the other BASELINE.json configs on one GPU (not the bench.py contract line;
evidence for SURVEY.md 8a rows a2-a6).  Cold L2 (256 MiB flush before each timed call), CUDA events, median.
Each line is JSON: config, ms per layer call, edges/s, algorithmic bytes (SURVEY.md 8d formula) and the
fraction of the measured HBM copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W

dev = torch.device("cuda", 0)
G.set_weight_cache(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6580.3
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, batch, D, ms, steps, extra_bytes_per_edge=0, extra_bytes=0):
    V, M, L = batch.num_nodes, batch.num_edges, len(batch.adjacency_lists)
    alg = steps * (M * (4 * D + 8 + extra_bytes_per_edge) + V * 8 * D + L * D * D * 4 + extra_bytes)
    print(json.dumps({"config": name, "V": V, "M": M, "L": L, "D": D, "timesteps_or_layers": steps,
                      "ms_per_call": round(ms, 4), "edges_per_s": M / (ms * 1e-3),
                      "algorithmic_bytes": alg, "achieved_GBps": alg / (ms * 1e-3) / 1e9,
                      "frac_of_measured_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak}), flush=True)


def states(V, D, seed=1):
    return torch.as_tensor(np.tanh(np.random.default_rng(seed).standard_normal((V, D))).astype(np.float32)).to(dev)


which = sys.argv[1:] or ["ggnn", "rgat", "film", "edge_mlp", "rgin", "rgcn5", "zipf"]
if "zipf" in which:   # degree-skewed variant of config 2 (Zipf(1.0) targets): hubs with thousands of incoming edges
    b = batching.ppi_like_batch(zipf_targets=True)
    deg = b.type_to_num_incoming_edges.sum(axis=0)
    h = states(b.num_nodes, 256)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(dev)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    w = W.to_torch(W.rgcn_weights(3, 256, 256), dev)
    ms = timeit(lambda: G.sparse_rgcn_layer(h, plan, cnt, 256, activation_function="ReLU", weights=w))
    report("RGCN PPI-shaped with Zipf(1.0) target skew (max in-degree %d, mean %.0f) hidden=256" % (int(deg.max()), float(deg.mean())),
           b, 256, ms, 1, extra_bytes_per_edge=4)
if "ggnn" in which:   # BASELINE config 3: GGNN QM9-shaped, 10k graphs, 4 bond types, hidden 128, 4 timesteps
    struct = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "qm9_valid_structure.npz")
    if os.path.exists(struct):   # the REAL 10,000 validation molecules (structure only), 4 bond types: V=180,560, M=373,466
        b, _, _ = batching.qm9_batch(batching.qm9_records_from_structure(struct), add_self_loop_edges=False)
    else:
        b = batching.qm9_like_batch(10000, seed=0)
    h = states(b.num_nodes, 128)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    w = W.to_torch(W.ggnn_weights(4, 128), dev)
    ms = timeit(lambda: G.sparse_ggnn_layer(h, plan, 128, num_timesteps=4, weights=w))
    report("GGNN QM9 10k graphs (V=%d M=%d, %s) hidden=128 4 timesteps (GRU)" % (b.num_nodes, b.num_edges, "real validation molecules" if os.path.exists(struct) else "synthetic"),
           b, 128, ms, 4, extra_bytes=0)
if "rgat" in which:   # config 4: RGAT PPI-shaped hidden 256, 8 heads
    b = batching.ppi_like_batch()
    h = states(b.num_nodes, 256)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    w = W.to_torch(W.rgat_weights(3, 256, 256), dev)
    ms = timeit(lambda: G.sparse_rgat_layer(h, plan, 256, num_heads=8, weights=w))
    report("RGAT PPI-shaped hidden=256 8 heads", b, 256, ms, 1, extra_bytes_per_edge=4 * 8)
if "film" in which:   # config 5 on ONE GPU: GNN-FiLM VarMisuse-shaped V=50k M=1M L=6 hidden 128
    b = batching.varmisuse_like_batch()
    h = states(b.num_nodes, 128)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(dev)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    w = W.to_torch(W.film_weights(6, 128, 128), dev)
    ms = timeit(lambda: G.sparse_gnn_film_layer(h, plan, cnt, 128, weights=w))
    report("GNN-FiLM VarMisuse-shaped V=50k M=1M L=6 hidden=128 (single GPU)", b, 128, ms, 1,
           extra_bytes=b.num_nodes * 6 * 8 * 128)
if "edge_mlp" in which:   # GNN-Edge-MLP1 on the PPI-shaped batch, hidden 256 (per-edge second Dense)
    b = batching.ppi_like_batch()
    h = states(b.num_nodes, 256)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(dev)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    for hid in (0, 1):
        w = W.to_torch(W.edge_mlp_weights(3, 256, 256, num_edge_hidden_layers=hid), dev)
        ms = timeit(lambda: G.sparse_gnn_edge_mlp_layer(h, plan, cnt, 256, activation_function="gelu",
                                                        num_edge_hidden_layers=hid, weights=w))
        report("GNN-Edge-MLP%d PPI-shaped hidden=256" % hid, b, 256, ms, 1, extra_bytes_per_edge=4 * 256 * (2 * hid))
if "rgin" in which:
    b = batching.ppi_like_batch()
    h = states(b.num_nodes, 256)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    w = W.to_torch(W.rgin_weights(3, 256, 256), dev)
    ms = timeit(lambda: G.sparse_rgin_layer(h, plan, 256, weights=w))
    report("RGIN PPI-shaped hidden=256 (edge MLP 1 hidden layer, source only)", b, 256, ms, 1)
if "rgcn5" in which:   # the packed 5-graph PPI batch (~max_nodes_in_batch 12,500)
    b = batching.ppi_like_batch(num_graphs=5)
    h = states(b.num_nodes, 256)
    cnt = torch.as_tensor(b.type_to_num_incoming_edges).to(dev)
    plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
    w = W.to_torch(W.rgcn_weights(3, 256, 256), dev)
    ms = timeit(lambda: G.sparse_rgcn_layer(h, plan, cnt, 256, activation_function="ReLU", weights=w))
    report("RGCN 5-graph packed PPI-shaped batch hidden=256", b, 256, ms, 1, extra_bytes_per_edge=4)
