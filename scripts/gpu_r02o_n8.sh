#!/bin/bash
# Round 2, N = 8, final code: the default line (extras included), then two grid caps of the overlapped transposes
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 2 > $OUT/bench_n8_$TAG.json 2> $OUT/bench_n8_$TAG.err
echo "bench exit $?"; tail -2 $OUT/bench_n8_$TAG.err; python scripts/show_bench.py $OUT/bench_n8_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_n8_$TAG.json").read().strip().splitlines()[-1])
print({k:v for k,v in d["nvlink"].items() if k!="note"})
for k,v in d.get("extra",{}).items():
    print(k, {q: v.get(q) for q in ("value","ms_per_step","check","unavailable","rel_l2_vs_oracle")})
    if isinstance(v,dict) and v.get("nvlink"): print("    ", {a:b for a,b in v["nvlink"].items() if a!="note"})
PY
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-extras --no-nccl-baseline > $OUT/bench_n8_${name}_$TAG.json 2> $OUT/bench_n8_${name}_$TAG.err
  echo "== $name ($*) exit $?"; python scripts/show_bench.py $OUT/bench_n8_${name}_$TAG.json | grep -E "value|poisson ms"
}
run s128 FEN_SLAB_SMS=128
run s112c2 FEN_SLAB_SMS=112 FEN_SLAB_CHUNKS=2
