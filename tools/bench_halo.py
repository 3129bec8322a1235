"""BASELINE config 5 on N GPUs: GNN-FiLM on ONE VarMisuse-shaped random graph (V=50k, M=1M, L=6, hidden 128),
node-range sharded with one all-to-all-v halo exchange of source rows per layer.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_halo.py

Prints one JSON line from rank 0: per-layer time (max over ranks, CUDA events) split into exchange and compute,
halo bytes, aggregate edges/s, and the max-norm difference of the sharded result from the single-GPU result."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching, weights as W
from tf_gnn_samples_b200.partition import NodeRangePartition

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
G.set_weight_cache(True)
D = 128
packed = int(os.environ.get("PACKED_GRAPHS", "0"))
b = batching.varmisuse_like_batch(packed_graphs=packed, seed=0)
h_all = np.tanh(np.random.default_rng(1).standard_normal((b.num_nodes, D))).astype(np.float32)
w = W.to_torch(W.film_weights(len(b.adjacency_lists), D, D), dev)
part = NodeRangePartition(b.adjacency_lists, b.type_to_num_incoming_edges, b.num_nodes, rank, world)
plan = G.GraphPlan(part.local_adjacency_lists, part.n_local, device=dev)
if os.environ.get("RESTRICT_TARGETS", "1") == "1":
    plan.set_num_targets(part.n_own)          # target-side GEMMs and the edge stage only for the owned rows
cnt = torch.as_tensor(part.local_num_incoming).to(dev)
h_own = torch.as_tensor(h_all[part.lo:part.hi]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def layer_from(local_states):
    return G.sparse_gnn_film_layer(local_states, plan, cnt, D, weights=w)[:part.n_own]


def timed(fn, n=30):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    t = torch.tensor([tot / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


full = timed(lambda: layer_from(part.exchange(h_own)))
exch = timed(lambda: part.exchange(h_own))
local_states = part.exchange(h_own)
comp = timed(lambda: layer_from(local_states))
out = layer_from(part.exchange(h_own))
# parity against the unsharded engine result (rank 0 computes the full graph on its GPU)
err = None
if world > 1:
    sizes = [None] * world
    dist.all_gather_object(sizes, int(out.shape[0]))
    gathered = [torch.empty((n, D), device=dev) for n in sizes]
    dist.all_gather(gathered, out.contiguous())
    if rank == 0:
        full_plan = G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev)
        ref = G.sparse_gnn_film_layer(torch.as_tensor(h_all).to(dev), full_plan, torch.as_tensor(b.type_to_num_incoming_edges).to(dev), D, weights=w)
        err = float((torch.cat(gathered) - ref).abs().max() / ref.abs().max())
halo = torch.tensor([part.n_halo], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "GNN-FiLM VarMisuse-shaped V=50k M=1M L=6 hidden=128, node-range sharded, %s" % ("packed %d graphs" % packed if packed else "one random graph"),
                      "n_gpus": world, "ms_per_layer": full, "ms_exchange": exch, "ms_compute": comp,
                      "edges_per_s": b.num_edges / (full * 1e-3), "max_halo_rows_per_rank": int(halo.item()),
                      "max_halo_bytes_per_rank_per_layer": int(halo.item()) * D * 4, "max_norm_diff_vs_single_gpu": err}))
if world > 1:
    dist.destroy_process_group()
