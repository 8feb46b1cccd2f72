#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_shortest_path.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_shortest_path.sh -input inputs/shortest_path -source 1
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.pregel.apps import main_submit_shortest_path as main; sys.exit(main(sys.argv[1:]))" "$@"
