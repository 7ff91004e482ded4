#!/usr/bin/env bash
# Multi-GPU visit for the paired-panel schedule: parity (dist_check) when CHECK=1, then the same factorisation
# under the schedules in SPECS (one NCCL communicator, one process group), then optionally the bench line.
set -u
TAG=$1; NG=$2; N3=$3; SPECS=$4; CHECK=${5:-0}; BENCH=${6:-0}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
if [ "$CHECK" = "1" ]; then
  timeout 300 $TR --master-port 29511 tools/dist_check.py 16384 2>&1 | grep "dist_check" | tee $OUT/${TAG}_dist_check_${NG}gpu.txt
fi
timeout 400 $TR --master-port 29517 tools/bench_configs_dist.py ${TAG}_${NG}gpu --n3 $N3 --schedules "$SPECS" 2>&1 | grep "^{" | cut -c1-400
if [ "$BENCH" = "1" ]; then
  timeout 500 $TR --master-port 29515 bench.py --gpus $NG --steps 2 --warmup 1 > $OUT/${TAG}_bench_${NG}gpu.json 2> $OUT/${TAG}_bench_${NG}gpu.err
  grep "^{" $OUT/${TAG}_bench_${NG}gpu.json | cut -c1-700; grep -v Warning $OUT/${TAG}_bench_${NG}gpu.err | tail -3
fi
