#!/bin/bash
# Elastic Table example `metric` (reference: services/et/bin/run_metric.sh)
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -m harmony_b200 et-example metric "$@"
