#!/usr/bin/env bash
# 8-GPU visit: the factorisation under the schedules in SPECS, then the bench line under the fastest of them
# (exported as environment; the code defaults are then set to it).
set -u
TAG=$1; NG=$2; N3=$3; SPECS=$4
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29517 tools/bench_configs_dist.py ${TAG}_${NG}gpu --n3 $N3 --schedules "$SPECS" 2>&1 | grep "^{" | cut -c1-400 | tee $OUT/${TAG}_schedules_${NG}gpu.txt
BEST=$(python - <<PY
import json
best=None
for l in open("$OUT/${TAG}_schedules_${NG}gpu.txt"):
    d=json.loads(l)
    if best is None or d["factor_ms"]<best[0]: best=(d["factor_ms"], d["schedule"])
print(best[1] if best else "AB_DIST_PAIR=1")
PY
)
echo "== bench under $BEST"
for kv in ${BEST//+/ }; do export "$kv"; done
timeout 500 $TR --master-port 29515 bench.py --gpus $NG --steps 2 --warmup 1 > $OUT/${TAG}_bench_${NG}gpu.json 2> $OUT/${TAG}_bench_${NG}gpu.err
grep "^{" $OUT/${TAG}_bench_${NG}gpu.json | cut -c1-900; grep -v Warning $OUT/${TAG}_bench_${NG}gpu.err | tail -3
echo "$BEST" > $OUT/${TAG}_bench_${NG}gpu.env
