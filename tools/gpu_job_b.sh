#!/bin/bash
# round-2 GPU job B: tests with the compact pair transform + PDL, per-config timings, launch lists, one full ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r02_gputests_c.log
python tools/bench_configs.py > $O/r02_configs_b.jsonl 2> $O/r02_configs_b.err
RGNN_NO_PAIRS=1 python tools/bench_configs.py ggnn > $O/r02_configs_b_nopairs.jsonl 2>> $O/r02_configs_b.err
python bench.py --steps 100 --warmup 5 --skip-cpu-baseline > $O/r02_bench_b_pdl.json 2> $O/r02_bench_b.err
RGNN_NO_PDL=1 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline > $O/r02_bench_b_nopdl.json 2>> $O/r02_bench_b.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches_ggnn_film.csv python tools/bench_configs.py ggnn film > /dev/null 2>> $O/r02_configs_b.err
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 6 -c 3 -o $O/r02_ggnn_gemm -f python tools/bench_configs.py ggnn > /dev/null 2>> $O/r02_configs_b.err
tail -4 $O/r02_gputests_c.log
cat $O/r02_configs_b.jsonl | cut -c1-260
