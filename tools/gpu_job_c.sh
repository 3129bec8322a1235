#!/bin/bash
# round-2 GPU job C (2 GPUs): regression check after the GATHER split, producer-group variants, the new bench line at N=1 and N=2
cd "$(dirname "$0")/.."
O=gpurun_out
export CUDA_VISIBLE_DEVICES=0
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r02_gputests_d.log
python tools/bench_configs.py > $O/r02_configs_c.jsonl 2> $O/r02_configs_c.err
for v in g3 g4; do RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_$v.so python tools/bench_configs.py ggnn film rgat rgcn5 > $O/r02_configs_c_$v.jsonl 2>> $O/r02_configs_c.err; done
python bench.py --steps 100 --warmup 5 > $O/r02_bench_c.json 2> $O/r02_bench_c.err
RGNN_NO_PDL=1 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs > $O/r02_bench_c_nopdl.json 2>> $O/r02_bench_c.err
for v in g3 g4; do RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_$v.so python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs --skip-e2e > $O/r02_bench_c_$v.json 2>> $O/r02_bench_c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-configs --skip-e2e > /dev/null 2>> $O/r02_bench_c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_film.csv python tools/bench_configs.py film > /dev/null 2>> $O/r02_configs_c.err
unset CUDA_VISIBLE_DEVICES
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > $O/r02_bench_c_n2.json 2> $O/r02_bench_c_n2.err
tail -3 $O/r02_gputests_d.log
cut -c1-200 $O/r02_configs_c.jsonl
tail -3 $O/r02_bench_c_n2.err
