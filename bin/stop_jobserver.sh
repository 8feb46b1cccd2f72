#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/stop_jobserver.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./stop_jobserver.sh -kill false
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.jobserver.client import stop_jobserver_main as main; sys.exit(main(sys.argv[1:]))" "$@"
