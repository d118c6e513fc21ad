#!/bin/bash
# N-GPU scaling round (one box): headline NS bench, Poisson-only bench, channel bench.  Usage: gpu_scale.sh tag N
TAG=$1; N=$2
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
PORT=29711
run() {  # name, extra args
    local name=$1; shift
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N "$@" > $OUT/${name}_${TAG}_n$N.json 2> $OUT/${name}_${TAG}_n$N.err
    echo "== $name N=$N exit $?"
    PORT=$((PORT+1))
    grep '^{' $OUT/${name}_${TAG}_n$N.json > $OUT/tmp.json
    python scripts/show_bench.py $OUT/tmp.json 2>/dev/null | grep -v "ghost_\|reduce  \|halo_unpack"
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/tmp.json").read().strip().splitlines()[-1])
    print("nvlink", d.get("nvlink")); print("check", d.get("check")); print("value", d.get("value"), d.get("unit"))
except Exception as e:
    print("no json", e)
PY
    grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" $OUT/${name}_${TAG}_n$N.err | tail -4
}
run bench --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2
run bench_poisson --steps 10 --warmup 3 --mode poisson
run bench_channel --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --case channel --size 512
