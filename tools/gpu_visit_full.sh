#!/usr/bin/env bash
# Round 2, GPU visit 1: parity suite, the per-config timings that round 1 never measured (configs 1, 4, 5), then the prepared kernel candidates (tools/sweep.sh: Gram k_r2_*, GEMM g_fast* and the TMA variant, potrf panel widths), then the bench line: everything the round needs from one box, most important first; each
# leg has its own timeout and writes to gpurun_out/ as it goes, so a cut-off call still leaves
# results.  Usage (under gpurun): bash tools/gpu_r2_visit1.sh [tag] [deadline_seconds]
set -u
# BEFORE calling gpurun: `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu` (same for tools/store_pattern) and
# `bash tools/sweep.sh build` here (3 min of CPU; the variant libraries travel with the
# snapshot).  Nothing is compiled on the GPU box.
TAG=${1:-r02a}
DEADLINE=${2:-1700}
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

leg 300 pytest bash -c "AB_RUN_UNVERIFIED=1 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -30 | tee $OUT/${TAG}_pytest.log"
leg 60 issue_probe bash -c "tools/issue_probe 2>&1 | tee $OUT/${TAG}_issue_probe.txt | tail -8"
leg 60 store_pattern bash -c "tools/store_pattern 32768 2>&1 | tee $OUT/${TAG}_store_pattern.txt | grep micro"
leg 600 configs bash -c "python tools/bench_configs.py ${TAG} 2>&1 | tail -12 | cut -c1-1500"
leg 400 sweep_gram env SWEEP_ONLY='k_*' SWEEP_TEST="tests/test_gpu_gram.py" bash tools/sweep.sh run ${TAG}_gram
leg 400 sweep_gemm env SWEEP_ONLY='g_*' bash tools/sweep.sh run ${TAG}_gemm
leg 200 sweep_potrf env SWEEP_ONLY='p_*' bash tools/sweep.sh run ${TAG}_potrf
leg 200 bench bash -c "python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 2500 $OUT/${TAG}_bench.json"
ls -la $OUT | tail -30
