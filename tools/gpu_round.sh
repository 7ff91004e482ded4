#!/usr/bin/env bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and the two full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1

echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee -a $OUT/${TAG}_pytest.log

echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json

echo "== ncu launch list (same bench command, shorter: 1 warm-up, 1 step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu \
    > $OUT/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
cat $OUT/${TAG}_launches_summary.txt

echo "== ncu full: Gram kernel (N=32768 SE+Matern52) and the DMMA GEMM (8192^3)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 2 \
    -f -o $OUT/${TAG}_gram python tools/first_light.py 32768 > $OUT/${TAG}_ncu_gram.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 2 \
    -f -o $OUT/${TAG}_gemm python tools/gemm_bench.py 1 > $OUT/${TAG}_ncu_gemm.log 2>&1
ls -la $OUT
