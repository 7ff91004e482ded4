#!/usr/bin/env bash
# Final multi-GPU visit: the bench line of this build at NG GPUs, then (optional) the factorisation under fewer NCCL channels.
set -u
TAG=$1; NG=$2; EXTRA=${3:-0}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29515 bench.py --gpus $NG --steps 2 --warmup 1 > $OUT/${TAG}_bench_${NG}gpu.json 2> $OUT/${TAG}_bench_${NG}gpu.err
grep "^{" $OUT/${TAG}_bench_${NG}gpu.json | cut -c1-600; grep -v Warning $OUT/${TAG}_bench_${NG}gpu.err | tail -3
if [ "$EXTRA" = "1" ]; then
  for ch in 8 4; do
    echo "== NCCL_MAX_NCHANNELS=$ch"
    NCCL_MAX_NCHANNELS=$ch timeout 200 $TR --master-port 29517 tools/bench_configs_dist.py ${TAG}_ch$ch --n3 131072 --schedules AB_DIST_NBUF=4 2>&1 | grep "^{" | cut -c1-300
  done
fi
