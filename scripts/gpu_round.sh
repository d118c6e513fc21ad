#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list and one full-set capture of a step.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_round.sh [tag] [quick]
TAG=${1:-r01}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log
tail -3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python scripts/show_bench.py $OUT/bench_$TAG.json


if [ -n "$QUICK" ]; then exit 0; fi
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline \
    > $OUT/bench_under_ncu_$TAG.log 2>&1
# full-set capture of one step's kernels
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_pred|k_rhs|k_corr|k_check|k_fft|k_thomas' \
    -s 40 -c 10 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline \
    > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT
