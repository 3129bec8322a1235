#!/bin/bash
# round-2 GPU job J (2 GPUs): LN split for small batches, GRU slab sizes, then bench --gpus 2 (coalesced .cg peer loads)
cd "$(dirname "$0")/.."
O=gpurun_out
export CUDA_VISIBLE_DEVICES=0
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02_gputests_j.log
python tools/bench_configs.py edge_mlp rgin > $O/r02_configs_j.jsonl 2> $O/r02_configs_j.err
RGNN_LN_SPLIT=0 python tools/bench_configs.py edge_mlp rgin > $O/r02_configs_j_nosplit.jsonl 2>> $O/r02_configs_j.err
for s in 37888 56832 75776 18944; do RGNN_GRU_SLAB=$s python tools/bench_configs.py ggnn | sed "s/^/slab=$s /" >> $O/r02_configs_j_slabs.txt 2>> $O/r02_configs_j.err; done
unset CUDA_VISIBLE_DEVICES
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 50 --warmup 5 > $O/r02_bench_j_n2.json 2> $O/r02_bench_j_n2.err
tail -3 $O/r02_gputests_j.log
for f in $O/r02_configs_j.jsonl $O/r02_configs_j_nosplit.jsonl $O/r02_configs_j_slabs.txt; do echo $f; cut -c1-190 $f; done
tail -2 $O/r02_bench_j_n2.err | cut -c1-200
