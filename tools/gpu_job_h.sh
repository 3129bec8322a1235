#!/bin/bash
# round-2 GPU job H (1 GPU): half-warp RGAT edge kernel, pair-GEMM subprocess test, full suite
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r02_gputests_h.log
python tools/bench_configs.py rgat > $O/r02_configs_h.jsonl 2> $O/r02_configs_h.err
RGNN_RGAT_HALF=0 python tools/bench_configs.py rgat > $O/r02_configs_h_nohalf.jsonl 2>> $O/r02_configs_h.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches_rgat_h.csv python tools/bench_configs.py rgat > /dev/null 2>> $O/r02_configs_h.err
tail -4 $O/r02_gputests_h.log; cut -c1-170 $O/r02_configs_h.jsonl $O/r02_configs_h_nohalf.jsonl; grep seg_rgat $O/r02_launches_rgat_h.csv | tail -3 | cut -c1-250
