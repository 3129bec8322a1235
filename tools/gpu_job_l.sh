#!/bin/bash
# round-2 GPU job L (1 GPU): slabbed edge stage + GRU, slabbed FiLM gamma/beta; smoke; A/B of both
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r02_gputests_l.log
python __graft_entry__.py --smoke > $O/r02_smoke_l.log 2>&1
python tools/bench_configs.py ggnn film > $O/r02_configs_l.jsonl 2> $O/r02_configs_l.err
RGNN_GRU_SLAB_EDGES_OFF=1 RGNN_FILM_SLAB_OFF=1 python tools/bench_configs.py ggnn film > $O/r02_configs_l_off.jsonl 2>> $O/r02_configs_l.err
tail -3 $O/r02_gputests_l.log; tail -3 $O/r02_smoke_l.log
for f in $O/r02_configs_l.jsonl $O/r02_configs_l_off.jsonl; do echo $f; cut -c1-190 $f; done
