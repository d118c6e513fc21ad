#!/bin/bash
# Round 2, N = 2: chunked overlap of the slab transposes: pieces 1 (none) / 4, grid caps
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -x -q --tb=short > $OUT/pytest_mp_$TAG.log 2>&1
echo "multiprocess tests exit $?"; tail -3 $OUT/pytest_mp_$TAG.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-extras --no-nccl-baseline > $OUT/bench_n2_${name}_$TAG.json 2> $OUT/bench_n2_${name}_$TAG.err
  echo "== $name ($*) exit $?"; python scripts/show_bench.py $OUT/bench_n2_${name}_$TAG.json | grep -E "value|poisson ms"
}
run c1 FEN_SLAB_CHUNKS=1
run c4 FEN_SLAB_CHUNKS=4
run c4s96 FEN_SLAB_CHUNKS=4 FEN_SLAB_SMS=96
run c4s32 FEN_SLAB_CHUNKS=4 FEN_SLAB_SMS=32
run c8 FEN_SLAB_CHUNKS=8
python scripts/show_bench.py $OUT/bench_n2_c4_$TAG.json
