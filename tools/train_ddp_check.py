"""Data-parallel training over graph-boundary shards on N GPUs (torchrun, NCCL): every rank trains RGCN/PPI on its shard of
one packed batch with the loss scaled by the GLOBAL node count and one gradient all-reduce per step; rank 0 also trains a
copy on the union batch alone.  After a few Adam steps the parameters must agree (fp32 rounding only), and the step
time is reported.  Usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_ddp_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import tf_gnn_samples_b200 as G
from tf_gnn_samples_b200 import batching
from tf_gnn_samples_b200.partition import split_batch_by_graphs
from tf_gnn_samples_b200.scaffold import RGCNPPIModel, global_count

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
saved_stdout = os.dup(1); os.dup2(2, 1)          # NCCL banners go to stderr; the JSON line is written to the saved fd

full = batching.pack_batch([batching.make_ppi_like_graph(2245, 59000, seed=40 + i) for i in range(2 * world)])
labels_full = (np.random.default_rng(7).random((full.num_nodes, 121)) < 0.3).astype(np.float32)
shards = split_batch_by_graphs(full, world)
lo = sum(s.num_nodes for s in shards[:rank])
mine = shards[rank]

def to_dev(b, labels):
    return (torch.as_tensor(b.node_features).to(dev), G.GraphPlan(b.adjacency_lists, b.num_nodes, device=dev),
            torch.as_tensor(b.type_to_num_incoming_edges).to(dev), torch.as_tensor(labels).to(dev))

# plain SGD for the comparison: Adam's first updates are ~lr * sign(g), which turns fp32 rounding differences of near-zero
# gradient entries into full-size parameter differences and would hide what is being checked (the summed gradient)
params = {"clamp_gradient_norm": 1.0, "learning_rate": 0.05, "random_seed": 0, "optimizer": "sgd"}
model = RGCNPPIModel(device=dev, params=params)
opt = model.make_optimizer()
f, p, c, y = to_dev(mine, labels_full[lo:lo + mine.num_nodes])
v_total = global_count(mine.num_nodes, dev)
STEPS = 5
for _ in range(STEPS):
    m = model.train_step_async(opt, f, p, c, y, global_num_nodes=v_total)
torch.cuda.synchronize()
# timing of the data-parallel step
ts = []
for _ in range(20):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); model.train_step_async(opt, f, p, c, y, global_num_nodes=v_total); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)

result = None
if rank == 0:
    ref = RGCNPPIModel(device=dev, params=params)
    ropt = ref.make_optimizer()
    ff, pp, cc, yy = to_dev(full, labels_full)
    model2 = RGCNPPIModel(device=dev, params=params)                     # fresh copy for the comparison run
    for _ in range(STEPS):
        ref.train_step_async(ropt, ff, pp, cc, yy)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
# compare after exactly STEPS steps: re-run the sharded training from scratch on every rank
model = RGCNPPIModel(device=dev, params=params); opt = model.make_optimizer()
for _ in range(STEPS):
    model.train_step_async(opt, f, p, c, y, global_num_nodes=v_total)
torch.cuda.synchronize()
if rank == 0:
    errs = {n: float((a.detach() - b.detach()).abs().max() / b.detach().abs().max())
            for (n, a), (_, b) in zip(model.named_parameters(), ref.named_parameters())}
    result = {"world": world, "graphs": 2 * world, "V": full.num_nodes, "M": full.num_edges, "steps_compared": STEPS,
              "max_param_rel_diff_vs_single_device": max(errs.values()), "worst": max(errs, key=errs.get),
              "data_parallel_step_ms": float(t.item()), "edges_per_s": full.num_edges / (float(t.item()) * 1e-3)}
    os.write(saved_stdout, (json.dumps(result) + "\n").encode())
if world > 1:
    dist.barrier(); dist.destroy_process_group()
