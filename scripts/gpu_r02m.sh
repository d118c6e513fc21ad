#!/bin/bash
# Round 2: the 32-values-per-thread z solve (fft_wide.cuh): parity at 512^3, then A/B against the radix-8 register path
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "512" > $OUT/pytest_parity_wide_$TAG.log 2>&1
echo "parity 512 (wide, 3 tiles) exit $?"; tail -4 $OUT/pytest_parity_wide_$TAG.log
FEN_SOLVE_WIDE=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "one_step_512" > $OUT/pytest_parity_wide2_$TAG.log 2>&1
echo "parity 512 (wide, 2 tiles) exit $?"; tail -4 $OUT/pytest_parity_wide2_$TAG.log
for V in 0 3 2; do
FEN_SOLVE_WIDE=$V timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_wide${V}_$TAG.json 2> $OUT/bench_wide${V}_$TAG.err
echo "FEN_SOLVE_WIDE=$V"; python scripts/show_bench.py $OUT/bench_wide${V}_$TAG.json | grep -E "value|fft_solve|poisson ms"
done
