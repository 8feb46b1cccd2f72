#!/bin/bash
# Elastic Table example `simple` (reference: services/et/bin/run_simple.sh)
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -m harmony_b200 et-example simple "$@"
