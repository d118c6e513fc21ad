#!/bin/bash
# Round 2: overlap correctness with 32 hardware queues (ranks as threads), c2r variants 6 / 7
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q --tb=short > $OUT/pytest_multirank_$TAG.log 2>&1
echo "multirank exit $?"; tail -5 $OUT/pytest_multirank_$TAG.log
FEN_SLAB_CHUNKS=8 timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q --tb=short -k "poisson_matches or 1024 or channel" > $OUT/pytest_multirank_c8_$TAG.log 2>&1
echo "chunks=8 exit $?"; tail -5 $OUT/pytest_multirank_c8_$TAG.log
FEN_X_C2R=7 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "tgv or poisson or 512" > $OUT/pytest_parity_c2rd_$TAG.log 2>&1
echo "parity with c2r_d+TWP exit $?"; tail -4 $OUT/pytest_parity_c2rd_$TAG.log
for V in 6 7; do
FEN_X_C2R=$V timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_c2r${V}_$TAG.json 2> $OUT/bench_c2r${V}_$TAG.err
echo "FEN_X_C2R=$V"; python scripts/show_bench.py $OUT/bench_c2r${V}_$TAG.json | grep -E "value|fft_x_c2r|poisson ms"
FEN_X_C2R=$V timeout 600 python bench.py --grid 1024,1024,128 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $OUT/bench_slab_c2r${V}_$TAG.json 2> $OUT/bench_slab_c2r${V}_$TAG.err
echo "slab 1024x1024x128 FEN_X_C2R=$V"; python scripts/show_bench.py $OUT/bench_slab_c2r${V}_$TAG.json | grep -E "fft_x_c2r"
done
