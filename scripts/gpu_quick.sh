#!/bin/bash
# quick single-GPU check: GPU tests + headline bench + optional extra bench commands (one per argument)
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -12 $OUT/pytest_gpu_$TAG.log
i=0
for CMD in "$@"; do
    i=$((i+1))
    echo "=== run $i: $CMD"
    timeout 900 bash -c "$CMD" > $OUT/run_${TAG}_$i.json 2> $OUT/run_${TAG}_$i.err
    grep '^{' $OUT/run_${TAG}_$i.json > $OUT/tmp.json
    python scripts/show_bench.py $OUT/tmp.json 2>/dev/null | grep -v "ghost_\|reduce  " || head -c 1500 $OUT/tmp.json
    tail -2 $OUT/run_${TAG}_$i.err
done
