#!/bin/bash
# A/B runs of the tuning switches (env vars read by libfen_gpu.so) on the bench; one block per variant.
# Usage: bash scripts/gpu_variants.sh tag "extra bench args" "VAR=val VAR2=val" "VAR=val" ...
#        (tag starting with nt_ skips the GPU test-suite)
TAG=$1; EXTRA=$2; shift; shift
OUT=gpurun_out
mkdir -p $OUT
if [[ $TAG != nt_* ]]; then
    timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
    echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
    tail -12 $OUT/pytest_gpu_$TAG.log
fi
i=0
for V in "$@"; do
    i=$((i+1))
    echo "=== variant $i: $V  [$EXTRA]"
    env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $EXTRA > $OUT/bench_${TAG}_v$i.json 2> $OUT/bench_${TAG}_v$i.err
    echo "# $V $EXTRA" >> $OUT/bench_${TAG}_v$i.json
    python scripts/show_bench.py $OUT/bench_${TAG}_v$i.json | grep -v "ghost\|reduce\|poisson ms"
    tail -2 $OUT/bench_${TAG}_v$i.err
done
