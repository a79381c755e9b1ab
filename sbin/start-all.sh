#!/usr/bin/env bash
# Master on this host + one worker per line of conf/workers (sbin/start-all.sh).  Lines are host[:gpu]; remote hosts are
# reached with ssh (the repo must be at the same path there), 127.0.0.1/localhost are started directly.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PORT="${AFB_MASTER_PORT:-7077}"; HOST="${AFB_MASTER_HOST:-127.0.0.1}"
"$ROOT/sbin/start-master.sh" --host "$HOST" --port "$PORT"
WORKERS="$ROOT/conf/workers"; [ -f "$WORKERS" ] || WORKERS="$ROOT/conf/workers.template"
i=0
grep -v '^\s*#' "$WORKERS" | grep -v '^\s*$' | while IFS=: read -r whost gpu; do
  args=("spark://$HOST:$PORT" --host "$whost" --cores "${AFB_WORKER_CORES:-2}" --memory "${AFB_WORKER_MEMORY:-4096}")
  [ -n "${gpu:-}" ] && args+=(--gpu "$gpu")
  if [ "$whost" = "127.0.0.1" ] || [ "$whost" = "localhost" ]; then
    AFB_WORKER_INSTANCE=$i "$ROOT/sbin/start-worker.sh" "${args[@]}"
  else
    ssh "$whost" "AFB_WORKER_INSTANCE=$i '$ROOT/sbin/start-worker.sh' ${args[*]}"
  fi
  i=$((i+1))
done
