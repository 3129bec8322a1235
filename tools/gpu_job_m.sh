#!/bin/bash
# round-2 GPU job M (1 GPU): final validation of the committed tree + the evidence files that go to profiles/
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02_gputests_final.log
python __graft_entry__.py --smoke > $O/r02_smoke_final.log 2>&1
python bench.py > $O/r02_bench_n1_final.json 2> $O/r02_bench_n1_final.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_arm.json 2>> $O/r02_bench_n1_final.err
python tools/bench_configs.py > $O/r02_configs_final.jsonl 2>> $O/r02_bench_n1_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-configs > /dev/null 2>> $O/r02_bench_n1_final.err
ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|seg_reduce' -s 12 -c 6 -o $O/r02_layer_final -f python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --skip-configs --skip-e2e > /dev/null 2>> $O/r02_bench_n1_final.err
python tools/ncu_traffic.py $O/r02_layer_final.ncu-rep $O/r02_traffic_final.json >> $O/r02_bench_n1_final.err 2>&1
tail -3 $O/r02_gputests_final.log; tail -2 $O/r02_smoke_final.log; cut -c1-400 $O/r02_bench_n1_final.json; echo; cut -c1-300 $O/r02_bench_reference_arm.json; echo; cut -c1-150 $O/r02_configs_final.jsonl
