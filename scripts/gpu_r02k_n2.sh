#!/bin/bash
# Round 2, N = 2: copy-engine transposes over real NVLink: IPC tests + timing against the bulk-store form
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651"
FEN_SLAB_DMA=1 FEN_SLAB_CHUNKS=4 timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -x -q --tb=short > $OUT/pytest_mp_dma_$TAG.log 2>&1
echo "multiprocess tests (copy engines) exit $?"; tail -3 $OUT/pytest_mp_dma_$TAG.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-extras --no-nccl-baseline > $OUT/bench_n2_${name}_$TAG.json 2> $OUT/bench_n2_${name}_$TAG.err
  echo "== $name ($*) exit $?"; python scripts/show_bench.py $OUT/bench_n2_${name}_$TAG.json | grep -E "value|poisson ms"
}
run bulk FEN_SLAB_DMA=0
run dma4 FEN_SLAB_DMA=1 FEN_SLAB_CHUNKS=4
run dma2 FEN_SLAB_DMA=1 FEN_SLAB_CHUNKS=2
run dma8 FEN_SLAB_DMA=1 FEN_SLAB_CHUNKS=8
python scripts/show_bench.py $OUT/bench_n2_dma4_$TAG.json; python - <<PY
import json
d=json.loads(open("$OUT/bench_n2_dma4_$TAG.json").read().strip().splitlines()[-1]); print({k:v for k,v in d["nvlink"].items() if k!="note"})
PY
