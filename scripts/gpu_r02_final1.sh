#!/bin/bash
# Round 2, one GPU: the whole GPU suite, smoke, the default bench line (with its extras), the reference arm, the
# ncu launch list and one full-set capture of the step's kernels, and the SASS evidence of the copy-engine instructions.
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log; tail -8 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log; tail -3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python scripts/show_bench.py $OUT/bench_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("cpu_baseline", d["cpu_baseline"]); print("roofline", d["roofline"])
for k,v in d.get("extra",{}).items():
    print(k, {q: v.get(q) for q in ("value","ms_per_step","unavailable")})
    for kk in (v.get("kernels") or [])[:14]: print("    ", kk.get("kernel"), kk.get("ms_per_step", kk.get("ms_per_solve")))
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
echo "reference arm exit $?"; python scripts/show_bench.py $OUT/bench_ref_$TAG.json | head -2
# launch list (cold-cache, serialised: shares only)
FEN_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 120 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras \
    > $OUT/bench_under_ncu_$TAG.log 2>&1
# full-set capture of one step's kernels
FEN_NO_GRAPH=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_pred|k_rhs|k_corr|k_check|k_fft|k_thomas' \
    -s 40 -c 10 -o $OUT/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras \
    > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT | tail -8
