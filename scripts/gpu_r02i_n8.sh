#!/bin/bash
# Round 2, N = 8: chunked overlap of the slab transposes: pieces 1 (none) / 4, grid caps 48 / 64 / 96 SMs
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-extras --no-nccl-baseline > $OUT/bench_n8_${name}_$TAG.json 2> $OUT/bench_n8_${name}_$TAG.err
  echo "== $name ($*) exit $?"; python scripts/show_bench.py $OUT/bench_n8_${name}_$TAG.json | grep -E "value|poisson ms"
}
run c1 FEN_SLAB_CHUNKS=1
run c4 FEN_SLAB_CHUNKS=4
run c4s96 FEN_SLAB_CHUNKS=4 FEN_SLAB_SMS=96
run c4s48 FEN_SLAB_CHUNKS=4 FEN_SLAB_SMS=48
python scripts/show_bench.py $OUT/bench_n8_c1_$TAG.json
