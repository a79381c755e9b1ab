#!/usr/bin/env bash
# Start a standalone master (sbin/start-master.sh of the reference).  Usage: sbin/start-master.sh [--host H] [--port P] [--webui-port W]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LOG_DIR="${AFB_LOG_DIR:-$ROOT/logs}"; mkdir -p "$LOG_DIR"
PYTHONPATH="$ROOT${PYTHONPATH:+:$PYTHONPATH}" nohup python -m asyncframework_b200.deploy.master "$@" \
  > "$LOG_DIR/master.out" 2>&1 &
echo $! > "$LOG_DIR/master.pid"
sleep 1; head -n 1 "$LOG_DIR/master.out" || true
