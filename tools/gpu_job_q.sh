#!/bin/bash
# round-2 GPU job Q (1 GPU): pack_b under programmatic dependent launch -- suite + uncached-weights number
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/r02_gputests_q.log
python bench.py --steps 100 --warmup 5 --skip-configs --skip-cpu-baseline --skip-e2e > $O/r02_bench_q.json 2> $O/r02_bench_q.err
tail -3 $O/r02_gputests_q.log; cut -c1-200 $O/r02_bench_q.json
