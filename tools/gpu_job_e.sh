#!/bin/bash
# round-2 GPU job E (1 GPU): heavy split + GRU slabs + producer-group guard; g3 variant; traffic capture
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r02_gputests_f.log
python tools/bench_configs.py zipf ggnn film > $O/r02_configs_e.jsonl 2> $O/r02_configs_e.err
RGNN_GRU_SLAB=1000000 python tools/bench_configs.py ggnn > $O/r02_configs_e_noslab.jsonl 2>> $O/r02_configs_e.err
RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_g3.so python -m pytest tests -m gpu -q -x -k "parity or golden or reference_pin or ops" 2>&1 | tail -6 > $O/r02_gputests_f_g3.log
RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_g3.so python tools/bench_configs.py > $O/r02_configs_e_g3.jsonl 2>> $O/r02_configs_e.err
RGNN_LIB_PATH=$PWD/tf-gnn-samples_b200/lib/librgnn_g3.so python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs --skip-e2e > $O/r02_bench_e_g3.json 2>> $O/r02_configs_e.err
python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-configs --skip-e2e > $O/r02_bench_e.json 2>> $O/r02_configs_e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_ggnn_e.csv python tools/bench_configs.py ggnn > /dev/null 2>> $O/r02_configs_e.err
ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|seg_reduce' -s 12 -c 6 -o $O/r02_layer -f python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-configs --skip-e2e > /dev/null 2>> $O/r02_configs_e.err
python tools/ncu_traffic.py $O/r02_layer.ncu-rep $O/r02_traffic.json >> $O/r02_configs_e.err 2>&1
tail -3 $O/r02_gputests_f.log; tail -3 $O/r02_gputests_f_g3.log
for f in $O/r02_configs_e.jsonl $O/r02_configs_e_noslab.jsonl $O/r02_configs_e_g3.jsonl; do echo $f; cut -c1-180 $f; done
