#!/bin/bash
# First GPU call of round 2: run what was written after round 1's GPU budget was spent, before anything else changes.
#   1. the whole GPU suite with the first-run files reported test by test (-rxX: XPASS = works, XFAIL = to fix);
#   2. smoke + headline bench (unchanged path: must reproduce profiles/r01v_bench.json);
#   3. A/B of the opt-in chunked host copies on the e2e figure (FEN_COPY_CHUNKS=4 vs default);
#   4. A/B of the persistent prefetching z-solve (FEN_FFT_SOLVE_PERSIST=1): the fft_solve row of the kernel table;
#   5. the any-length Poisson path on a 384^3 grid (3 x 2^7 in every direction) next to 512^3: ms/step and kernels.
#   6. compute-sanitizer memcheck of the never-run kernels.
# Usage (repo root, on the GPU box):  bash scripts/gpu_r02_first.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxX > $OUT/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
tail -40 $OUT/pytest_gpu_$TAG.log
# the first-run files once more with their xfail marks ignored: full tracebacks of whatever does not pass yet
timeout 1500 python -m pytest tests/test_gpu_zy_any_length.py tests/test_gpu_zz_rising_bubble.py -q --runxfail \
    > $OUT/pytest_firstrun_$TAG.log 2>&1
echo "first-run files exit $?"; tail -30 $OUT/pytest_firstrun_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1
echo "smoke exit $?" >> $OUT/smoke_$TAG.log; tail -3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench exit $?"; python scripts/show_bench.py $OUT/bench_$TAG.json
FEN_COPY_CHUNKS=4 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_chunks4_$TAG.json 2> $OUT/bench_chunks4_$TAG.err
FEN_COPY_CHUNKS=8 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_chunks8_$TAG.json 2> $OUT/bench_chunks8_$TAG.err
python - <<PY
import json
for k in ("", "_chunks4", "_chunks8"):
    try:
        d = json.loads(open("gpurun_out/bench%s_%s.json" % (k, "$TAG")).read().strip().splitlines()[-1])
        print("e2e%s: %.0f Mcell-updates/s, %.1f ms/step" % (k, d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    except Exception as exc:
        print("e2e%s: no line (%r)" % (k, exc))
PY
FEN_FFT_SOLVE_PERSIST=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_persist_$TAG.json 2> $OUT/bench_persist_$TAG.err
FEN_X_C2R=4 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_c2rp_$TAG.json 2> $OUT/bench_c2rp_$TAG.err
echo "persistent c2r exit $?"; python scripts/show_bench.py $OUT/bench_c2rp_$TAG.json 2>/dev/null | head -12
echo "persistent fft_solve exit"; python scripts/show_bench.py $OUT/bench_persist_$TAG.json 2>/dev/null | head -12
timeout 600 python bench.py --grid 384,384,384 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_any384_$TAG.json 2> $OUT/bench_any384_$TAG.err
echo "any-length 384^3 exit $?"; python scripts/show_bench.py $OUT/bench_any384_$TAG.json 2>/dev/null | head -20
# memcheck of the kernels that have never run: the any-length suite and the new operators under compute-sanitizer
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zy_any_length.py -q --runxfail \
    -k "poisson_any_length or scalar_laplacian or cavity_48x40" > $OUT/memcheck_any_$TAG.log 2>&1
echo "memcheck exit $?"; tail -5 $OUT/memcheck_any_$TAG.log
ls -la $OUT | tail -12
