#!/usr/bin/env bash
# Run the CPU test-suite while N busy-loop processes compete for the cores (docs/TESTING.md: timing-based assertions must hold on
# a loaded machine).  Usage: scripts/run_cpu_suite_under_load.sh [N_BUSY=6] [extra pytest args...]
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
N="${1:-6}"; shift || true
pids=()
for _ in $(seq 1 "$N"); do
  python -c "
import time
t = time.time()
while time.time() - t < 1500: sum(i * i for i in range(10000))
" &
  pids+=($!)
done
trap 'kill "${pids[@]}" 2>/dev/null' EXIT
cd "$ROOT" && python -m pytest tests/ -q -m "not gpu" -p no:cacheprovider --tb=short "$@"
