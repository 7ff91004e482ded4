#!/usr/bin/env bash
# Round 2, GPU visit b: the whole GPU suite on the new defaults (achieved errors logged), the bench line, the
# ncu launch list of the bench command and full captures of the two headline kernels.
set -u
TAG=${1:-r02b}
DEADLINE=${2:-1100}
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
left() { echo $(( DEADLINE - ($(date +%s) - T0) )); }
leg() {
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
leg 420 pytest bash -c "AB_ERR_LOG=$PWD/$OUT/${TAG}_achieved_errors.tsv python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -60 | tee $OUT/${TAG}_pytest.log"
leg 200 bench bash -c "python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 1500 $OUT/${TAG}_bench.json"
leg 150 ncu_gram bash -c "ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -f -o $OUT/${TAG}_gram python tools/gram_bench.py 32768 7 3 2 > $OUT/${TAG}_ncu_gram.log 2>&1; tail -3 $OUT/${TAG}_ncu_gram.log"
leg 150 ncu_gemm bash -c "ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tma -c 1 -f -o $OUT/${TAG}_gemm python tools/gemm_bench.py 8 > $OUT/${TAG}_ncu_gemm.log 2>&1; tail -3 $OUT/${TAG}_ncu_gemm.log"
leg 300 ncu_launches bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $OUT/${TAG}_launches_n32768.csv python bench.py --n 32768 --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1; python tools/summarize_launches.py $OUT/${TAG}_launches_n32768.csv | tee $OUT/${TAG}_launches_n32768.txt | head -20; rm -f $OUT/${TAG}_launches_n32768.csv"
ls -la $OUT | tail -20
