#!/bin/bash
# Full GPU test-suite + the wave2d bench line (no ncu).  Usage: bash scripts/gpu_suite.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -25 $OUT/pytest_gpu_$TAG.log
timeout 300 python bench.py --case wave2d --steps 20 --warmup 6 > $OUT/bench_wave2d_$TAG.json 2> $OUT/bench_wave2d_$TAG.err
echo "bench exit $?"; tail -3 $OUT/bench_wave2d_$TAG.err; python scripts/show_bench.py $OUT/bench_wave2d_$TAG.json
FEN_THOMAS_LP=0 timeout 300 python bench.py --case wave2d --steps 10 --warmup 6 --no-e2e > $OUT/bench_wave2d_nolp_$TAG.json 2>> $OUT/bench_wave2d_$TAG.err
python scripts/show_bench.py $OUT/bench_wave2d_nolp_$TAG.json | head -8
