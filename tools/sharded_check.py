"""BASELINE config 5 on N GPUs through the library's own sharded path (rgnn_halo_plan_create / rgnn_halo_exchange over
CUDA-IPC peer memory): GNN-FiLM on ONE VarMisuse-shaped graph (V=50k, M=1M, L=6, hidden 128), node-range sharded.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py [--packed G] [--layers K]

Rank 0 prints one JSON line: parity of the reassembled result against the reference-generated fixture
(tests/golden/ref_config5_film_*.npz: the reference's own gnn_film.py through tests/tf1_shim) for ONE layer, per-layer time of a
K-layer stack (CUDA events, max over ranks, CUDA-graph replay), the exchange kernel alone, halo bytes and NVLink GB/s."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch, torch.distributed as dist
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W

ap = argparse.ArgumentParser()
ap.add_argument("--packed", type=int, default=0)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
G.set_weight_cache(True)
D = 128
b = batching.varmisuse_like_batch(packed_graphs=args.packed, seed=0)
h_all = np.tanh(np.random.default_rng(1).standard_normal((b.num_nodes, D))).astype(np.float32)
ws = [W.to_torch(W.film_weights(len(b.adjacency_lists), D, D, seed=2 + 10 * i), dev) for i in range(args.layers)]
cuts = G.degree_balanced_cuts(b.adjacency_lists, b.num_nodes, world)
sg = G.ShardedGraph(b.adjacency_lists, cuts, rank, world, device=dev)
sg.attach(D)
cnt = sg.local_num_incoming(b.type_to_num_incoming_edges)
h_own = torch.as_tensor(h_all[sg.lo:sg.hi]).to(dev)


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])


def stack(layers):
    for t in range(layers):
        sg.exchange(t % 2)
        G.sparse_gnn_film_layer(sg.states(t % 2), sg.plan, cnt, D, weights=ws[t], out=sg.states(1 - t % 2))


# ---- parity: ONE layer against the reference-generated fixture (whole graph) ----
sg.states(0)[: sg.n_own] = h_own
torch.cuda.synchronize(); barrier()
stack(1)
torch.cuda.synchronize(); barrier()
mine = sg.states(1)[: sg.n_own].contiguous()
parity = None
if world > 1:
    sizes = [None] * world
    dist.all_gather_object(sizes, int(mine.shape[0]))
    parts = [torch.empty((n, D), device=dev) for n in sizes]
    dist.all_gather(parts, mine)
    full = torch.cat(parts)
else:
    full = mine
if rank == 0:
    import ref_cases as RC
    z = np.load(RC.fixture_path("config5_film_packed" if args.packed else "config5_film_random"))
    if args.packed in (0, 25):
        parity = dict(zip(("rows", "projection", "colsum"), RC.compare_with_summary(full.cpu().numpy(), z)))
        parity["reference_float32_path"] = float(z["err32"])


# ---- timing ----
def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize(); barrier()
    t = torch.tensor([s.elapsed_time(e) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


K = args.layers - args.layers % 2          # an even number of layers returns to buffer 0: a replayable step
eager_ms = timed(lambda: stack(K), args.iters) / K
exch_ms = timed(lambda: (sg.exchange(0), sg.exchange(1)), args.iters) / 2
graph_ms = None
try:
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        stack(K)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(); barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        stack(K)
    torch.cuda.synchronize(); barrier()
    graph_ms = timed(g.replay, args.iters) / K
except Exception as exc:
    print("graph capture failed on rank %d: %r" % (rank, exc), file=sys.stderr)
halo = torch.tensor([sg.n_halo, sg.n_local, sg.plan.num_edges], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
if rank == 0:
    best = graph_ms if graph_ms is not None else eager_ms
    hb = int(halo[0]) * D * 4
    print(json.dumps({"config": "GNN-FiLM VarMisuse-shaped V=50k M=1M L=6 hidden=128, node-range sharded (librgnn peer-pull exchange), %s"
                      % ("packed %d graphs" % args.packed if args.packed else "one random graph"),
                      "n_gpus": world, "layers": K, "ms_per_layer_graph": graph_ms, "ms_per_layer_eager": eager_ms,
                      "ms_exchange_kernel": exch_ms, "edges_per_s": b.num_edges / (best * 1e-3),
                      "max_halo_rows_per_rank": int(halo[0]), "max_local_rows_per_rank": int(halo[1]), "max_local_edges": int(halo[2]),
                      "halo_bytes_per_rank_per_layer": hb, "exchange_GBps_per_rank": hb / (exch_ms * 1e-3) / 1e9 if exch_ms > 0 else None,
                      "parity_vs_reference_fixture_one_layer": parity}), flush=True)
sg.close()
if world > 1:
    dist.destroy_process_group()
