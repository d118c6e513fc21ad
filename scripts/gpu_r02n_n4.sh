#!/bin/bash
# Round 2, N = 4 (1024 x 1024 x 512): first run of the default line on 4 real GPUs (4 pieces, 96-SM cap, 512-point z lines)
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29671"
timeout 900 $TR bench.py --gpus 4 --steps 10 --warmup 3 --e2e-steps 2 > $OUT/bench_n4_$TAG.json 2> $OUT/bench_n4_$TAG.err
echo "bench exit $?"; tail -2 $OUT/bench_n4_$TAG.err; python scripts/show_bench.py $OUT/bench_n4_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_n4_$TAG.json").read().strip().splitlines()[-1])
print({k:v for k,v in d["nvlink"].items() if k!="note"})
for k,v in d.get("extra",{}).items():
    print(k, {q: v.get(q) for q in ("value","ms_per_step","check","unavailable","rel_l2_vs_oracle")})
PY
FEN_SLAB_CHUNKS=1 timeout 400 $TR bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e --no-extras --no-nccl-baseline > $OUT/bench_n4_c1_$TAG.json 2> $OUT/bench_n4_c1_$TAG.err
echo "== unchunked exit $?"; python scripts/show_bench.py $OUT/bench_n4_c1_$TAG.json | grep -E "value|poisson ms"
