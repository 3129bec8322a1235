#!/bin/bash
# round-2 GPU job D (8 GPUs): tests on GPU 0 (heavy split), then bench.py --gpus 8 with the sharded config-5 block
cd "$(dirname "$0")/.."
O=gpurun_out
CUDA_VISIBLE_DEVICES=0 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > $O/r02_gputests_e.log
CUDA_VISIBLE_DEVICES=0 python tools/bench_configs.py zipf > $O/r02_configs_d_zipf.jsonl 2> $O/r02_configs_d.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 50 --warmup 5 > $O/r02_bench_d_n8.json 2> $O/r02_bench_d_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 50 --warmup 5 > $O/r02_bench_d_n4.json 2> $O/r02_bench_d_n4.err
tail -3 $O/r02_gputests_e.log; cat $O/r02_configs_d_zipf.jsonl | cut -c1-200; tail -3 $O/r02_bench_d_n8.err | cut -c1-300
