#!/usr/bin/env bash
# GPU visit 2: parity tests (look-ahead potrf, leaf v2, C++ layer), bench line, store-pattern probe,
# tuning sweep, ncu full captures.  Usage (under gpurun): bash tools/gpu_round2.sh [tag]
set -u
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee -a $OUT/${TAG}_pytest.log
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json
echo "== store pattern probe"
timeout 300 tools/store_pattern 32768 2>&1 | tee $OUT/${TAG}_store_pattern.txt
echo "== sweep"
SWEEP_SKIP="g_128x128x16s4 g_64x128x16s4c2 k_cols1" bash tools/sweep.sh run $TAG
echo "== ncu full: Gram kernel (N=32768 SE+Matern52) and the DMMA GEMM (8192^3)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 \
    -f -o $OUT/${TAG}_gram python tools/gram_bench.py 32768 7 3 3 > $OUT/${TAG}_ncu_gram.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 \
    -f -o $OUT/${TAG}_gemm python tools/gemm_bench.py 1 > $OUT/${TAG}_ncu_gemm.log 2>&1
ls -la $OUT
