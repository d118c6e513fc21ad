#!/bin/bash
# Round-end evidence in one gpurun call: full GPU test-suite, smoke, headline bench + reference arm, wave2d bench,
# ncu launch lists and full-set captures of both paths.   Usage: bash scripts/gpu_final.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
bash scripts/gpu_round.sh $TAG
timeout 300 python bench.py --case wave2d --steps 20 --warmup 6 > $OUT/bench_wave2d_$TAG.json 2> $OUT/bench_wave2d_$TAG.err
echo "wave2d bench exit $?"; python scripts/show_bench.py $OUT/bench_wave2d_$TAG.json 2>/dev/null | head -20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 100 --csv \
    --log-file $OUT/launches_wave2d_$TAG.csv python bench.py --case wave2d --steps 4 --warmup 6 --no-e2e \
    > $OUT/bench_under_ncu_wave2d_$TAG.log 2>&1
FEN_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vof|k_mf|k_thomas_lp' \
    -s 40 -c 10 -o $OUT/prof_wave2d_$TAG -f python bench.py --case wave2d --steps 2 --warmup 6 --no-e2e \
    > $OUT/ncu_full_wave2d_$TAG.log 2>&1
ls -la $OUT | tail -20
