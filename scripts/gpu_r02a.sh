#!/bin/bash
# Round 2, first GPU call: tracebacks of the five hidden two-phase failures + A/B of the opt-in switches.
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_zz_rising_bubble.py tests/test_gpu_zy_any_length.py -q --runxfail --tb=long -rA \
    > $OUT/pytest_firstrun_$TAG.log 2>&1
echo "first-run files exit $?"; tail -15 $OUT/pytest_firstrun_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python scripts/show_bench.py $OUT/bench_$TAG.json
for K in 4 8; do
FEN_COPY_CHUNKS=$K timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_chunks${K}_$TAG.json 2> $OUT/bench_chunks${K}_$TAG.err
done
python - <<PY
import json
for k in ("", "_chunks4", "_chunks8"):
    try:
        d = json.loads(open("gpurun_out/bench%s_%s.json" % (k, "$TAG")).read().strip().splitlines()[-1])
        print("e2e%s: %.0f Mcell-updates/s, %.1f ms/step" % (k, d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"].get("cpu_affinity"))
    except Exception as exc:
        print("e2e%s: no line (%r)" % (k, exc))
PY
FEN_FFT_SOLVE_PERSIST=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_persist_$TAG.json 2> $OUT/bench_persist_$TAG.err
FEN_X_C2R=4 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_c2rp_$TAG.json 2> $OUT/bench_c2rp_$TAG.err
echo "persistent c2r"; python scripts/show_bench.py $OUT/bench_c2rp_$TAG.json 2>/dev/null | head -14
echo "persistent fft_solve"; python scripts/show_bench.py $OUT/bench_persist_$TAG.json 2>/dev/null | head -14
