#!/usr/bin/env bash
# 1-GPU: GPU suite after the paired-panel change, then world-1 runs of the distributed factorisation with and
# without pairing at nb = 512 / 1024 (what the k-depth of the bulk update is worth without any communication).
set -u
TAG=${1:-r02r}
OUT=gpurun_out; mkdir -p $OUT
leg() { local max=$1 name=$2; shift 2; echo "== $name (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; timeout "$max" "$@"; echo "   rc=$? (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; }
leg 600 pytest bash -c "AB_ERR_LOG=$PWD/$OUT/${TAG}_achieved_errors.tsv python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -60 | tee $OUT/${TAG}_pytest.log"
leg 300 w1 bash -c "python tools/dist_w1_bench.py 65536 AB_DIST_NB=512 AB_DIST_NB=512+AB_DIST_PAIR=0 AB_DIST_NB=1024 AB_DIST_NB=1024+AB_DIST_PAIR=0 AB_DIST_NB=512+AB_DIST_NBUF=8 AB_DIST_NB=2048+AB_DIST_PAIR=0 2>&1 | tee $OUT/${TAG}_dist_w1.txt"
