#!/usr/bin/env bash
# Final 1-GPU visit of the round: the bench line of this build, then the GPU suite.
set -u
TAG=${1:-r02final}
OUT=gpurun_out; mkdir -p $OUT
timeout 170 python bench.py > $OUT/${TAG}_bench_n65536.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$? (t+$SECONDS)"
grep "^{" $OUT/${TAG}_bench_n65536.json | cut -c1-400
AB_ERR_LOG=$PWD/$OUT/${TAG}_achieved_errors.tsv timeout 170 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.log; echo "pytest done (t+$SECONDS)"
