#!/usr/bin/env bash
# Start a worker and register it with a master (sbin/start-slave.sh).  Usage: sbin/start-worker.sh spark://host:7077 [--cores N] [--memory MB] [--gpu I]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LOG_DIR="${AFB_LOG_DIR:-$ROOT/logs}"; mkdir -p "$LOG_DIR"
N="${AFB_WORKER_INSTANCE:-0}"
PYTHONPATH="$ROOT${PYTHONPATH:+:$PYTHONPATH}" nohup python -m asyncframework_b200.deploy.worker "$@" \
  > "$LOG_DIR/worker-$N.out" 2>&1 &
echo $! > "$LOG_DIR/worker-$N.pid"
