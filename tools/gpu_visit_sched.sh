#!/usr/bin/env bash
# Schedule comparison of the distributed factorisation in ONE torchrun: bash tools/gpu_visit_sched.sh <tag> <ngpus> <n> <schedules>
set -u
TAG=$1; NG=$2; N3=$3; SCHED=$4
mkdir -p gpurun_out
timeout ${5:-500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/bench_configs_dist.py $TAG --n3 $N3 --schedules "$SCHED" 2>&1 | grep -v Warning | grep "^{" | cut -c1-400
if [ "${TUNE_BENCH:-0}" = "1" ]; then
  timeout 200 python tools/tune_bench.py $TAG --n 16384 --evals 96 --gpus $NG 2>&1 | tail -2 | cut -c1-1500
fi
