#!/bin/bash
# round-2 GPU job P (1 GPU): epoch-loop tests against the real models + the whole suite once more
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_training_gpu.py -m gpu -q 2>&1 | tail -30 > $O/r02_gputests_training.log
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r02_gputests_p.log
cat $O/r02_gputests_training.log | tail -25; tail -3 $O/r02_gputests_p.log
