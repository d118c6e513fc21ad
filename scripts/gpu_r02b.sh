#!/bin/bash
# Round 2, second call: the blocked slab path (bulk-store transposes) on one device + the repaired two-phase tests + 512^3 parity
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q --tb=short > $OUT/pytest_multirank_$TAG.log 2>&1
echo "multirank exit $?"; tail -15 $OUT/pytest_multirank_$TAG.log
timeout 1500 python -m pytest tests/test_gpu_zz_rising_bubble.py "tests/test_gpu_parity.py::test_one_step_512_matches_c_oracle" -m gpu -q --tb=short > $OUT/pytest_fixed_$TAG.log 2>&1
echo "fixed tests exit $?"; tail -15 $OUT/pytest_fixed_$TAG.log
