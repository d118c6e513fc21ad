#!/bin/bash
# Round 2: copy-engine transposes, correctness on one device (ranks as threads), all forms
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q --tb=short > $OUT/pytest_multirank_$TAG.log 2>&1
echo "multirank exit $?"; tail -12 $OUT/pytest_multirank_$TAG.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "staged_x or one_step_tgv3d" > $OUT/pytest_parity_$TAG.log 2>&1
echo "parity exit $?"; tail -4 $OUT/pytest_parity_$TAG.log
