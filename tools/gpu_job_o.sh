#!/bin/bash
# round-2 GPU job O (1 GPU): last validation of the committed tree
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02_gputests_last.log
python __graft_entry__.py --smoke 2>&1 | tail -2 > $O/r02_smoke_last.log
python bench.py --steps 100 --warmup 5 --skip-configs --skip-cpu-baseline > $O/r02_bench_last.json 2> $O/r02_bench_last.err
tail -3 $O/r02_gputests_last.log; cat $O/r02_smoke_last.log; cut -c1-250 $O/r02_bench_last.json
