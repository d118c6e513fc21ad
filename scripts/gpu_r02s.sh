#!/bin/bash
# Round 2: vof sweeps with every face flux evaluated once: two-phase parity + the wave2d bench
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multiphase.py tests/test_gpu_zz_rising_bubble.py -m gpu -x -q --tb=short -k "not lid3d and not reference_resolution and not shear_drop" > $OUT/pytest_mf_$TAG.log 2>&1
echo "two-phase tests exit $?"; tail -4 $OUT/pytest_mf_$TAG.log
timeout 600 python bench.py --case wave2d --steps 40 --warmup 6 --no-e2e --no-cpu-baseline > $OUT/bench_wave2d_$TAG.json 2> $OUT/bench_wave2d_$TAG.err
echo "wave2d exit $?"; python scripts/show_bench.py $OUT/bench_wave2d_$TAG.json | head -20
