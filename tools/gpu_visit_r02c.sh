#!/usr/bin/env bash
# Round 2, GPU visit c (1 GPU): the GPU suite on the fused one-rhs solves, solve-phase timing, bench line.
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
leg() { local max=$1 name=$2; shift 2; echo "== $name" | tee -a $OUT/${TAG}_legs.txt; timeout "$max" "$@"; echo "   rc=$?" | tee -a $OUT/${TAG}_legs.txt; }
leg 600 pytest bash -c "AB_ERR_LOG=$PWD/$OUT/${TAG}_achieved_errors.tsv python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log"
leg 120 potrf bash -c "python tools/potrf_bench.py 32768 2>&1 | tee $OUT/${TAG}_potrf.txt; python tools/potrf_bench.py 65536 2 2>&1 | tee -a $OUT/${TAG}_potrf.txt"
leg 60 gemv bash -c "python tools/gemm_bench.py 19 2>&1 | tail -3 | tee $OUT/${TAG}_gemv.txt"
leg 200 bench bash -c "python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1200 $OUT/${TAG}_bench.json"
