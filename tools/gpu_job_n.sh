#!/bin/bash
# round-2 GPU job N (2 GPUs): bench --gpus 2 with the overlapped halo exchange
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 50 --warmup 5 > $O/r02_bench_n_n2.json 2> $O/r02_bench_n_n2.err
tail -2 $O/r02_bench_n_n2.err | cut -c1-200; cut -c1-200 $O/r02_bench_n_n2.json
