#!/bin/bash
# round-2 GPU job K (8 GPUs): bench.py --gpus 8 with the committed sharded path (coalesced .cg peer loads)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 8 --steps 50 --warmup 5 > $O/r02_bench_k_n8.json 2> $O/r02_bench_k_n8.err
tail -2 $O/r02_bench_k_n8.err | cut -c1-200; cut -c1-300 $O/r02_bench_k_n8.json
