#!/bin/bash
# Round 2, final code on one GPU: the whole GPU suite (with the isotropic-turbulence driver) and smoke
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -s -k "isotropic" > $OUT/pytest_isotropic_$TAG.log 2>&1
echo "isotropic exit $?"; grep -E "isotropic 128|passed|failed|Error" $OUT/pytest_isotropic_$TAG.log | tail -5
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_zx_isotropic.py > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log; tail -6 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log; tail -3 $OUT/smoke_$TAG.log
