#!/usr/bin/env bash
# 1-GPU: GPU suite (new tests), world-1 distributed factorisation vs potrf, ncu of the final GEMM, launch list.
set -u
TAG=${1:-r02p}
OUT=gpurun_out; mkdir -p $OUT
leg() { local max=$1 name=$2; shift 2; echo "== $name (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; timeout "$max" "$@"; echo "   rc=$? (t+$SECONDS)" | tee -a $OUT/${TAG}_legs.txt; }
leg 600 pytest bash -c "AB_ERR_LOG=$PWD/$OUT/${TAG}_achieved_errors.tsv python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log"
leg 200 w1 bash -c "python tools/dist_w1_bench.py 65536 2>&1 | tee $OUT/${TAG}_dist_w1.txt"
leg 150 ncu_gemm bash -c "ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tma -c 1 -f -o $OUT/${TAG}_gemm python tools/gemm_bench.py 1 > $OUT/${TAG}_ncu_gemm.log 2>&1; tail -2 $OUT/${TAG}_ncu_gemm.log"
leg 200 ncu_launches bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $OUT/${TAG}_launches_n16384.csv python bench.py --n 16384 --steps 1 --warmup 1 --no-cpu --no-extras > $OUT/${TAG}_ncu_bench.log 2>&1; python tools/summarize_launches.py $OUT/${TAG}_launches_n16384.csv | tee $OUT/${TAG}_launches_n16384.txt | head -16; rm -f $OUT/${TAG}_launches_n16384.csv"
