#!/bin/bash
# Round 2, N = 8: the default bench line (headline + Poisson-only + channel + parity extras) and the 8-GPU IPC test
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 2 > $OUT/bench_n8_$TAG.json 2> $OUT/bench_n8_$TAG.err
echo "bench exit $?"; tail -3 $OUT/bench_n8_$TAG.err; python scripts/show_bench.py $OUT/bench_n8_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_n8_$TAG.json").read().strip().splitlines()[-1])
print(json.dumps(d["nvlink"])[:1500])
print("e2e", d["e2e"] and d["e2e"]["value"], "check", d["check"])
for k,v in d.get("extra",{}).items():
    print(k, {q: v.get(q) for q in ("value","ms_per_step","poisson_solve_ms","check","unavailable","rel_l2_vs_oracle")}, json.dumps(v.get("nvlink"))[:600] if isinstance(v,dict) else "")
    for kk in (v.get("kernels") or [])[:9]: print("    ", kk.get("kernel"), kk.get("ms_per_step", kk.get("ms_per_solve")))
PY
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -x -q --tb=short > $OUT/pytest_mp_$TAG.log 2>&1
echo "multiprocess tests exit $?"; tail -5 $OUT/pytest_mp_$TAG.log
