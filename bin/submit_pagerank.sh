#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_pagerank.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_pagerank.sh -input inputs/adj_list
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.pregel.apps import main_submit_pagerank as main; sys.exit(main(sys.argv[1:]))" "$@"
