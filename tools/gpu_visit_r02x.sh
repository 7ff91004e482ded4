#!/usr/bin/env bash
# 1-GPU: distributed tests + world-1 A/B of the split column update (diagonal block first, rows below on a second
# panel stream while the block is factored).
set -u
TAG=${1:-r02x}
OUT=gpurun_out; mkdir -p $OUT
leg() { local max=$1 name=$2; shift 2; echo "== $name (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; timeout "$max" "$@"; echo "   rc=$? (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; }
leg 300 pytest_dist bash -c "python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_dist.log"
leg 300 w1 bash -c "python tools/dist_w1_bench.py 65536 AB_DIST_NB=512 AB_DIST_NB=512+AB_DIST_SPLITCOL=0 AB_DIST_NB=1024 AB_DIST_NB=1024+AB_DIST_SPLITCOL=0 2>&1 | tee $OUT/${TAG}_dist_w1.txt"
