#!/bin/bash
# round-2 GPU job I (1 GPU): x32 TMEM epilogue loads (ld16 variant for A/B), cheaper RGAT softmax update, TMA bulk row-gather experiment
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02_gputests_i.log
python tools/bench_configs.py > $O/r02_configs_i.jsonl 2> $O/r02_configs_i.err
RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_ld16.so python tools/bench_configs.py ggnn film rgcn5 > $O/r02_configs_i_ld16.jsonl 2>> $O/r02_configs_i.err
python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs --skip-e2e > $O/r02_bench_i.json 2>> $O/r02_configs_i.err
RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_ld16.so python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs --skip-e2e > $O/r02_bench_i_ld16.json 2>> $O/r02_configs_i.err
RGNN_SEG_BULK=1 python -m pytest tests -m gpu -q -x -k "rgcn or golden" 2>&1 | tail -4 > $O/r02_gputests_i_bulk.log
RGNN_SEG_BULK=1 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs --skip-e2e > $O/r02_bench_i_bulk.json 2>> $O/r02_configs_i.err
RGNN_SEG_BULK=1 python tools/bench_configs.py rgcn5 > $O/r02_configs_i_bulk.jsonl 2>> $O/r02_configs_i.err
RGNN_SEG_BULK=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_bulk.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-configs --skip-e2e > /dev/null 2>> $O/r02_configs_i.err
tail -3 $O/r02_gputests_i.log; tail -3 $O/r02_gputests_i_bulk.log
for f in $O/r02_configs_i.jsonl $O/r02_configs_i_ld16.jsonl $O/r02_configs_i_bulk.jsonl; do echo $f; cut -c1-170 $f; done
for f in $O/r02_bench_i.json $O/r02_bench_i_ld16.json $O/r02_bench_i_bulk.json; do cut -c1-160 $f; done
grep seg_reduce_bulk $O/r02_launches_bulk.csv | tail -2 | cut -c1-260
