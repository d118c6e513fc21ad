#!/bin/bash
# 8-GPU channel bench (BASELINE configs[2] shape 1024x1024x512, ppn): staggered scatter vs fused back-substitution stores
TAG=${1:-r01}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
for MODE in scatter fused; do
  if [ $MODE = fused ]; then export FEN_THOMAS_FUSED_A2A=1; else unset FEN_THOMAS_FUSED_A2A; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
      bench.py --gpus $N --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --case channel --size 512 \
      > $OUT/bench_channel_${MODE}_${TAG}_n$N.json 2> $OUT/bench_channel_${MODE}_${TAG}_n$N.err
  echo "== channel $MODE N=$N exit $?"
  grep '^{' $OUT/bench_channel_${MODE}_${TAG}_n$N.json > $OUT/tmp.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/tmp.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "nvlink", d.get("nvlink"))
    for k in d["kernels"][:10]: print("  ", k["kernel"], round(k["ms_per_step"], 3))
except Exception as e:
    print("no json", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" $OUT/bench_channel_${MODE}_${TAG}_n$N.err | tail -3
done
