#!/bin/bash
# Round 2: row-private r2c (FEN_X_R2C=5 / 6) and twiddle products in the strided passes (libfen_gpu_twp.so)
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
FEN_X_R2C=6 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "tgv or poisson or 512" > $OUT/pytest_parity_r2c6_$TAG.log 2>&1
echo "parity with r2c_v+TWP exit $?"; tail -4 $OUT/pytest_parity_r2c6_$TAG.log
FEN_GPU_LIB=$PWD/fen_b200/libfen_gpu_twp.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -x -q --tb=short -k "tgv or poisson or 512" > $OUT/pytest_parity_twp_$TAG.log 2>&1
echo "parity with strided TWP exit $?"; tail -4 $OUT/pytest_parity_twp_$TAG.log
for V in 1 5 6; do
FEN_X_R2C=$V timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_r2c${V}_$TAG.json 2> $OUT/bench_r2c${V}_$TAG.err
echo "FEN_X_R2C=$V"; python scripts/show_bench.py $OUT/bench_r2c${V}_$TAG.json | grep -E "value|fft_x_r2c|poisson ms"
done
for V in 1 6; do
FEN_X_R2C=$V timeout 600 python bench.py --grid 1024,1024,128 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_slab_r2c${V}_$TAG.json 2> $OUT/bench_slab_r2c${V}_$TAG.err
echo "slab 1024x1024x128 FEN_X_R2C=$V"; python scripts/show_bench.py $OUT/bench_slab_r2c${V}_$TAG.json | grep -E "fft_x_r2c"
done
FEN_GPU_LIB=$PWD/fen_b200/libfen_gpu_twp.so timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_twp_$TAG.json 2> $OUT/bench_twp_$TAG.err
echo "strided TWP 512^3"; python scripts/show_bench.py $OUT/bench_twp_$TAG.json | grep -E "value|fft_|poisson ms"
FEN_GPU_LIB=$PWD/fen_b200/libfen_gpu_twp.so timeout 600 python bench.py --grid 1024,1024,128 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_slab_twp_$TAG.json 2> $OUT/bench_slab_twp_$TAG.err
echo "strided TWP slab"; python scripts/show_bench.py $OUT/bench_slab_twp_$TAG.json | grep -E "value|fft_|poisson ms"
