#!/bin/bash
# One gpurun call for the two-phase path: its GPU parity tests, the wave2d bench line, a launch list and one
# full-set ncu capture of its kernels.   Usage: bash scripts/gpu_mf.sh [tag]
TAG=${1:-r01mf}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multiphase.py -m gpu -q -x --tb=short > $OUT/pytest_mf_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_mf_$TAG.log
tail -40 $OUT/pytest_mf_$TAG.log
timeout 600 python -m pytest tests/test_gpu_multiphase.py -m gpu -q --tb=line > $OUT/pytest_mf_all_$TAG.log 2>&1
tail -15 $OUT/pytest_mf_all_$TAG.log
timeout 300 python bench.py --case wave2d --steps 20 --warmup 6 > $OUT/bench_wave2d_$TAG.json 2> $OUT/bench_wave2d_$TAG.err
echo "bench exit $?"; tail -3 $OUT/bench_wave2d_$TAG.err; python scripts/show_bench.py $OUT/bench_wave2d_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv \
    --log-file $OUT/launches_wave2d_$TAG.csv python bench.py --case wave2d --steps 4 --warmup 6 --no-e2e \
    > $OUT/bench_under_ncu_wave2d_$TAG.log 2>&1
FEN_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vof|k_mf' \
    -s 40 -c 8 -o $OUT/prof_wave2d_$TAG -f python bench.py --case wave2d --steps 2 --warmup 6 --no-e2e \
    > $OUT/ncu_full_wave2d_$TAG.log 2>&1
ls -la $OUT | tail -12
