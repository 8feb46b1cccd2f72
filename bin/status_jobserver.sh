#!/bin/bash
# Running / finished jobs of the job server (STATUS command on the command port; not in the reference).
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
exec python -m harmony_b200 status "$@"
