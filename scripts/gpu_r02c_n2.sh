#!/bin/bash
# Round 2, N = 2: bulk-store transposes against the register-store epilogues, the IPC tests, and the full default line
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -x -q --tb=short > $OUT/pytest_mp_$TAG.log 2>&1
echo "multiprocess tests exit $?"; tail -5 $OUT/pytest_mp_$TAG.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-extras > $OUT/bench_n2_bulk_$TAG.json 2> $OUT/bench_n2_bulk_$TAG.err
echo "bulk exit $?"; python scripts/show_bench.py $OUT/bench_n2_bulk_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_n2_bulk_$TAG.json").read().strip().splitlines()[-1]); print(json.dumps(d["nvlink"])[:900])
PY
FEN_SLAB_BULK=0 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-extras --no-nccl-baseline > $OUT/bench_n2_regs_$TAG.json 2> $OUT/bench_n2_regs_$TAG.err
echo "register-store exit $?"; python scripts/show_bench.py $OUT/bench_n2_regs_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_n2_regs_$TAG.json").read().strip().splitlines()[-1]); print(json.dumps(d["nvlink"])[:600])
PY
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 3 > $OUT/bench_n2_full_$TAG.json 2> $OUT/bench_n2_full_$TAG.err
echo "full line exit $?"; tail -3 $OUT/bench_n2_full_$TAG.err; python - <<PY
import json
d=json.loads(open("$OUT/bench_n2_full_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"] and d["e2e"]["value"], "check", d["check"])
for k,v in d.get("extra",{}).items(): print(k, json.dumps(v)[:700])
PY
