#!/usr/bin/env bash
# Development aid: builds differently-tuned variants of the library (only gemm.o / gram_fixed.o are
# recompiled, with -D overrides) into tools/sweep/ and, with "run", times each on the GPU box.
#   bash tools/sweep.sh build        (here, no GPU)
#   bash tools/sweep.sh run [tag]    (under gpurun)
set -u
cd "$(dirname "$0")/.."
CS=albatross_b200/csrc
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DAB_BUILDING"
OUT=tools/sweep
# tag : source : defines
VARIANTS=(
  "k_base:gram_fixed:"
  "k_r1_nocheck:gram_fixed:-DAB_GRAM_ONECHECK=0"
  "p_nb2048:linalg:"
  "p_nb1024:linalg:-DAB_POTRF_NB=1024"
  "p_nb4096:linalg:-DAB_POTRF_NB=4096"
  "g_base:gemm:"
  "g_nofast:gemm:-DAB_GEMM_FASTLOAD=0"
  "g_fast_s4:gemm:-DAB_GEMM_STAGES=4"
)
SKIP_RUN="${SWEEP_SKIP:-}"
if [ "${1:-build}" = "build" ]; then
  make -s -j8 -C $CS || exit 1
  mkdir -p $OUT
  for v in "${VARIANTS[@]}"; do
    IFS=: read -r tag src defs <<< "$v"
    if [ -n "${SWEEP_ONLY:-}" ]; then case "$tag" in ${SWEEP_ONLY}) ;; *) continue;; esac; fi
    (
      objs=$(ls $CS/build/*.o)
      mine=""
      for one in ${src//,/ }; do
        $NV $defs -c -o $OUT/${tag}_$one.o $CS/$one.cu || exit 1
        objs=$(echo "$objs" | grep -v "/$one.o")
        mine="$mine $OUT/${tag}_$one.o"
      done
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$tag.so $mine $objs -cudart static -ldl
      rm -f $mine
      echo "built $tag"
    ) &
  done
  wait
  ls -la $OUT
  exit 0
fi
TAG=${2:-sweep}
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_sweep.txt
: > $LOG
for v in "${VARIANTS[@]}"; do
  IFS=: read -r tag src defs <<< "$v"
  case " $SKIP_RUN " in *" $tag "*) continue;; esac
  if [ -n "${SWEEP_ONLY:-}" ]; then case "$tag" in ${SWEEP_ONLY}) ;; *) continue;; esac; fi
  echo "== $tag ($defs)" | tee -a $LOG
  src=${src%%,*}
  if [ "$src" = "linalg" ]; then
    ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/potrf_bench.py 32768 2>&1 | tee -a $LOG
    ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/potrf_bench.py 65536 2 2>&1 | tee -a $LOG
    if [ "$tag" = "p_nb2048" ]; then
      AB_POTRF_RECURSIVE=1 ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/potrf_bench.py 65536 2 2>&1 | sed 's/^/recursive: /' | tee -a $LOG
    fi
  elif [ "$src" = "gemm" ]; then
    if [ "$tag" = "g_base" ]; then
      # the cp.async kernel alone (the TMA kernel of the NT product is on by default)
      AB_GEMM_TMA=0 ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 120 python tools/gemm_bench.py 8 2>&1 | sed 's/^/no-tma: /' | tee -a $LOG
      AB_GEMM_TMA=0 ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 120 python tools/potrf_bench.py 32768 2>&1 | sed 's/^/no-tma: /' | tee -a $LOG
    fi
    ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/gemm_bench.py 8 2>&1 | tee -a $LOG
    ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/potrf_bench.py 32768 2>&1 | tee -a $LOG
  else
    if [ -n "${SWEEP_TEST:-}" ]; then
      ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 200 python -m pytest $SWEEP_TEST -m gpu -q -x 2>&1 | tail -3 | tee -a $LOG
    fi
    ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/gram_bench.py 32768 7 3 6 2>&1 | tail -1 | tee -a $LOG
    ALBATROSS_B200_LIB=$PWD/$OUT/lib_$tag.so timeout 300 python tools/gram_bench.py 32768 7 3 6 1 2>&1 | tail -1 | tee -a $LOG
  fi
done
