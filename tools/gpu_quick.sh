#!/usr/bin/env bash
# Short GPU visit: run the given pytest selection (default: everything marked gpu) and optional extra command.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-quick}
SEL=${2:-tests}
shift 2 || true
timeout 1200 python -m pytest $SEL -m gpu -x -q 2>&1 | tail -40 | tee $OUT/${TAG}_pytest.log
if [ $# -gt 0 ]; then
  echo "== extra: $*"
  timeout 900 "$@" 2>&1 | tail -60 | tee $OUT/${TAG}_extra.log
fi
