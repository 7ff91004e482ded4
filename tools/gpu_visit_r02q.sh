#!/usr/bin/env bash
# 1-GPU: GPU suite (polynomial tests, sub-blocked distributed solves), then the A/B of reading C into the
# accumulators at the start of the TMA GEMM (AB_GEMM_CINIT=0 restores the epilogue read): kernel, potrf, world-1 dist.
set -u
TAG=${1:-r02q}
OUT=gpurun_out; mkdir -p $OUT
leg() { local max=$1 name=$2; shift 2; echo "== $name (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; timeout "$max" "$@"; echo "   rc=$? (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; }
leg 600 pytest bash -c "AB_ERR_LOG=$PWD/$OUT/${TAG}_achieved_errors.tsv python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log"
leg 100 gemm_cinit1 bash -c "python tools/gemm_bench.py 10 2>&1 | sed 's/^/cinit=1 /' | tee $OUT/${TAG}_gemm_ab.txt"
leg 100 gemm_cinit0 bash -c "AB_GEMM_CINIT=0 python tools/gemm_bench.py 10 2>&1 | sed 's/^/cinit=0 /' | tee -a $OUT/${TAG}_gemm_ab.txt"
leg 100 potrf1 bash -c "python tools/potrf_bench.py 65536 2 2>&1 | sed 's/^/cinit=1 /' | tee $OUT/${TAG}_potrf_ab.txt"
leg 100 potrf0 bash -c "AB_GEMM_CINIT=0 python tools/potrf_bench.py 65536 2 2>&1 | sed 's/^/cinit=0 /' | tee -a $OUT/${TAG}_potrf_ab.txt"
leg 200 w1 bash -c "python tools/dist_w1_bench.py 65536 AB_DIST_NB=1024 AB_DIST_NB=512 AB_DIST_NB=2048 2>&1 | tee $OUT/${TAG}_dist_w1.txt"
