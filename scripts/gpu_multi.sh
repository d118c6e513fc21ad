#!/bin/bash
# Multi-GPU round on one box: IPC multi-process parity test, NS bench and Poisson-only bench at N GPUs.
# Usage: bash scripts/gpu_multi.sh tag N [with_tests]
TAG=$1; N=$2; TESTS=${3:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu_$TAG.txt 2>&1
if [ -n "$TESTS" ]; then
    timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
    echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
    tail -12 $OUT/pytest_gpu_$TAG.log
else
    timeout 900 python -m pytest tests/test_gpu_multiprocess.py -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1
    echo "pytest exit $?" >> $OUT/pytest_gpu_$TAG.log
    tail -12 $OUT/pytest_gpu_$TAG.log
fi
PORT=29611
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err
echo "bench N=$N exit $?"; grep '^{' $OUT/bench_${TAG}_n$N.json > $OUT/tmp.json; python scripts/show_bench.py $OUT/tmp.json
python - <<PY
import json
d=json.loads(open("$OUT/tmp.json").read().strip().splitlines()[-1])
print("nvlink", d.get("nvlink"))
PY
tail -3 $OUT/bench_${TAG}_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --gpus $N --steps 10 --warmup 3 --mode poisson > $OUT/bench_poisson_${TAG}_n$N.json 2> $OUT/bench_poisson_${TAG}_n$N.err
echo "poisson bench N=$N exit $?"; grep '^{' $OUT/bench_poisson_${TAG}_n$N.json | cut -c1-1500
tail -3 $OUT/bench_poisson_${TAG}_n$N.err
timeout 600 python bench.py --steps 10 --warmup 3 --mode poisson > $OUT/bench_poisson_${TAG}_n1.json 2> $OUT/bench_poisson_${TAG}_n1.err
echo "poisson bench N=1 exit $?"; grep '^{' $OUT/bench_poisson_${TAG}_n1.json | cut -c1-1500
