#!/bin/bash
# Runs the EXAMPLE USAGE line of every run_*.sh (reference: jobserver/bin/run_all.sh:16-25)
cd "$(dirname "${BASH_SOURCE[0]}")"
PYTHONPATH="$(cd .. && pwd):$PYTHONPATH" python -m harmony_b200.common.samples . > /dev/null   # sample_* inputs (idempotent)
for f in run_*.sh; do
  [ "$f" = "run_all.sh" ] && continue
  cmd=$(grep -A1 "EXAMPLE USAGE" "$f" | tail -1 | sed 's/^# //')
  [ -z "$cmd" ] && cmd="./$f"
  echo "== $cmd"; eval "$cmd" || echo "FAILED: $f"
done
