#!/bin/bash
# round-2 GPU job F (1 GPU): 3 producer groups as default (g2 variant for A/B), fused RGAT scores, launch lists
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02_gputests_g.log
python tools/bench_configs.py > $O/r02_configs_f.jsonl 2> $O/r02_configs_f.err
RGNN_RGAT_UNFUSED=1 python tools/bench_configs.py rgat > $O/r02_configs_f_rgat_unfused.jsonl 2>> $O/r02_configs_f.err
RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_g2.so python tools/bench_configs.py ggnn film rgcn5 > $O/r02_configs_f_g2.jsonl 2>> $O/r02_configs_f.err
python bench.py --steps 100 --warmup 5 > $O/r02_bench_f.json 2> $O/r02_bench_f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r02_launches_rgat_edgemlp.csv python tools/bench_configs.py rgat edge_mlp > /dev/null 2>> $O/r02_configs_f.err
tail -4 $O/r02_gputests_g.log
for f in $O/r02_configs_f.jsonl $O/r02_configs_f_rgat_unfused.jsonl $O/r02_configs_f_g2.jsonl; do echo $f; cut -c1-170 $f; done
