#!/usr/bin/env bash
# Multi-GPU visit: bash tools/gpu_visit_dist.sh <tag> <ngpus> [n3] [only]
set -u
TAG=${1:-r02e}
NG=${2:-2}
N3=${3:-65536}
ONLY=${4:-3,4,5}
OUT=gpurun_out
mkdir -p $OUT
leg() { local max=$1 name=$2; shift 2; echo "== $name (t+$SECONDS s)" | tee -a $OUT/${TAG}_legs.txt; timeout "$max" "$@"; echo "   rc=$? (t+$SECONDS s)" | tee -a $OUT/${TAG}_legs.txt; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > $OUT/${TAG}_gpu.txt 2>&1
[ "${SKIP_PYTEST:-0}" = "1" ] || leg 200 pytest_dist bash -c "python -m pytest tests/test_gpu_dist.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log"
leg 300 dist_check bash -c "$TR --master-port 29511 tools/dist_check.py 16384 2>&1 | grep -v Warning | tail -40 | tee $OUT/${TAG}_dist_check.txt"
for n3 in ${N3//,/ }; do
  leg 600 configs_$n3 bash -c "$TR --master-port 29513 tools/bench_configs_dist.py $TAG --only $ONLY --n3 $n3 2>&1 | grep -v Warning | tail -12 | cut -c1-2500"
  ONLY=3
done
if [ "${BENCH_STEPS:-0}" != "0" ]; then
  leg 600 bench bash -c "$TR --master-port 29515 bench.py --gpus $NG --steps $BENCH_STEPS --warmup ${BENCH_WARMUP:-1} > $OUT/${TAG}_bench_${NG}gpu.json 2> $OUT/${TAG}_bench_${NG}gpu.err; tail -c 3000 $OUT/${TAG}_bench_${NG}gpu.json; grep -v Warning $OUT/${TAG}_bench_${NG}gpu.err | tail -5"
fi
