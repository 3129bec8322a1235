#!/bin/bash
# round-2 GPU job G (1 GPU): CTA pairs (tcgen05 cta_group::2) -- correctness first (bounded waits trap instead of hanging), then timing
cd "$(dirname "$0")/.."
O=gpurun_out
RGNN_GEMM_PAIR=1 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -25 > $O/r02_pair_tests_ops.log
if grep -q "passed" $O/r02_pair_tests_ops.log && ! grep -q "failed" $O/r02_pair_tests_ops.log; then
  RGNN_GEMM_PAIR=1 timeout 600 python -m pytest tests -m gpu -q -x -k "parity or golden or reference_pin or models" 2>&1 | tail -15 > $O/r02_pair_tests_all.log
  RGNN_GEMM_PAIR=1 python tools/bench_configs.py ggnn film rgcn5 edge_mlp > $O/r02_configs_g_pair.jsonl 2> $O/r02_configs_g.err
fi
tail -12 $O/r02_pair_tests_ops.log; tail -5 $O/r02_pair_tests_all.log 2>/dev/null; cut -c1-170 $O/r02_configs_g_pair.jsonl 2>/dev/null
