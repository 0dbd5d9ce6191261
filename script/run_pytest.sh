#!/usr/bin/env bash
# script/run_pytest.sh cpu|gpu|multigpu [extra pytest args]   (counterpart of the reference's script/run_pytest.sh)
set -euo pipefail
cd "$(dirname "$0")/.."
kind="${1:-cpu}"; shift || true
case "$kind" in
  cpu) python -m pytest tests -x -q -m "not gpu" "$@";;
  gpu) python -m pytest tests -x -q -m gpu "$@";;
  multigpu) python -m pytest tests/test_multigpu.py -x -q -m gpu "$@";;
  *) echo "usage: $0 cpu|gpu|multigpu"; exit 2;;
esac
