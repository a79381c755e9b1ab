#!/usr/bin/env bash
# Stop the daemons started by start-all.sh on this host (by recorded PID, never by pattern).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LOG_DIR="${AFB_LOG_DIR:-$ROOT/logs}"
for f in "$LOG_DIR"/worker-*.pid "$LOG_DIR"/master.pid; do
  [ -f "$f" ] || continue
  pid="$(cat "$f")"; kill "$pid" 2>/dev/null || true; rm -f "$f"
done
