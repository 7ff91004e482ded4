#!/usr/bin/env bash
# GPU visit 4 (last of round 1: conflict-free Gram table + swizzled staging, 2-CTA GEMM tiles, nb=2048 look-ahead): everything the round needs from one box, most important first; each
# leg has its own timeout and writes to gpurun_out/ as it goes, so a cut-off call still leaves
# results.  Usage (under gpurun): bash tools/gpu_round4.sh [tag] [deadline_seconds]
set -u
TAG=${1:-r01d}
DEADLINE=${2:-840}
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
left() { echo $(( DEADLINE - ($(date +%s) - T0) )); }
leg() { # leg <max seconds> <name> <command...>: skipped when the deadline is too close
  local max=$1 name=$2
  shift 2
  local l
  l=$(left)
  if [ "$l" -lt 20 ]; then
    echo "== skip $name (deadline)" | tee -a $OUT/${TAG}_legs.txt
    return
  fi
  [ "$max" -gt "$l" ] && max=$l
  echo "== $name (t+$(( $(date +%s) - T0 )) s, limit $max s)" | tee -a $OUT/${TAG}_legs.txt
  timeout "$max" "$@"
  echo "   rc=$? (t+$(( $(date +%s) - T0 )) s)" | tee -a $OUT/${TAG}_legs.txt
}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1

leg 300 pytest bash -c "python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log"
leg 120 sweep_gram env SWEEP_ONLY='k_*' bash tools/sweep.sh run ${TAG}_gram
leg 200 sweep_potrf env SWEEP_ONLY='p_*' bash tools/sweep.sh run ${TAG}_potrf
leg 200 bench bash -c "python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench.json"
leg 100 ncu_gram bash -c "ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 \
    -f -o $OUT/${TAG}_gram python tools/gram_bench.py 32768 7 3 3 > $OUT/${TAG}_ncu_gram.log 2>&1"
leg 200 bench_nb4096 bash -c "ALBATROSS_B200_LIB=$PWD/tools/sweep/lib_p_nb4096.so python bench.py --no-cpu > $OUT/${TAG}_bench_nb4096.json 2> $OUT/${TAG}_bench_nb4096.err; tail -c 1200 $OUT/${TAG}_bench_nb4096.json"
leg 200 sweep_gemm env SWEEP_ONLY='g_*' bash tools/sweep.sh run ${TAG}_gemm
leg 100 ncu_gemm bash -c "ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 \
    -f -o $OUT/${TAG}_gemm python tools/gemm_bench.py 1 > $OUT/${TAG}_ncu_gemm.log 2>&1"
leg 200 ncu_launches bash -c "ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches_n32768.csv python bench.py --n 32768 --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1; \
    python tools/summarize_launches.py $OUT/${TAG}_launches_n32768.csv > $OUT/${TAG}_launches_n32768.txt 2>&1; head -20 $OUT/${TAG}_launches_n32768.txt"
ls -la $OUT | tail -30
