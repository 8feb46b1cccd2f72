#!/bin/bash
# Elastic Table example `simple` (reference: services/et/bin/run_simple.sh)
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
exec python -c "import sys; from harmony_b200.et.examples import main; sys.exit(main(['simple']))"
