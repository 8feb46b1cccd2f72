#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/start_jobserver.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./start_jobserver.sh -num_executors 8 -scheduler edu.snu.cay.jobserver.driver.SchedulerImpl -executor_num_tasklets 4 -timeout 0 -backend auto
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.jobserver.client import start_jobserver_main as main; sys.exit(main(sys.argv[1:]))" "$@"
