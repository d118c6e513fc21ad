#!/bin/bash
# A/B runs of the tuning switches (env vars read by libfen_gpu.so) on the 512^3 bench; one line per variant.
# Usage: bash scripts/gpu_variants.sh tag "VAR=val VAR2=val" "VAR=val" ...
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -12 $OUT/pytest_gpu_$TAG.log
i=0
for V in "$@"; do
    i=$((i+1))
    echo "=== variant $i: $V"
    env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_${TAG}_v$i.json 2> $OUT/bench_${TAG}_v$i.err
    echo "# $V" >> $OUT/bench_${TAG}_v$i.json
    python scripts/show_bench.py $OUT/bench_${TAG}_v$i.json | grep -v "ghost\|reduce\|poisson ms"
    tail -2 $OUT/bench_${TAG}_v$i.err
done
