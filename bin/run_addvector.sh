#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/run_addvector.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./run_addvector.sh -num_dolphin_workers 2 -num_servers 2 -max_num_epochs 3 -num_mini_batches 8 -num_worker_blocks 8 -num_keys 10 -vector_size 4 -delta 1 -compute_time_ms 50 -num_training_data 64
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.examples import main_run_addvector as main; sys.exit(main(sys.argv[1:]))" "$@"
