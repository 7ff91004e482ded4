#!/usr/bin/env bash
# GPU visit 5 (last minutes of round 1: instruction-count variants of the Gram kernel, each parity-tested before it is timed; final bench line): everything the round needs from one box, most important first; each
# leg has its own timeout and writes to gpurun_out/ as it goes, so a cut-off call still leaves
# results.  Usage (under gpurun): bash tools/gpu_round5.sh [tag] [deadline_seconds]
set -u
TAG=${1:-r01e}
DEADLINE=${2:-330}
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

leg 110 gram_all3 env SWEEP_ONLY='k_all3' SWEEP_TEST="tests/test_gpu_gram.py tests/test_gpu_fullsize.py tests/test_gpu_dist.py tests/test_gpu_sparse.py" bash tools/sweep.sh run ${TAG}_all3
leg 130 gram_singles env SWEEP_ONLY='k_[bpoec]*' SWEEP_TEST="tests/test_gpu_gram.py" bash tools/sweep.sh run ${TAG}_singles
leg 100 bench bash -c "python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 900 $OUT/${TAG}_bench.json"
leg 100 bench_all3 bash -c "ALBATROSS_B200_LIB=$PWD/tools/sweep/lib_k_all3.so python bench.py --no-cpu > $OUT/${TAG}_bench_all3.json 2> $OUT/${TAG}_bench_all3.err; tail -c 900 $OUT/${TAG}_bench_all3.json"
ls -la $OUT | tail -12
