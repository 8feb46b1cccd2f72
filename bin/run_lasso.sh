#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/run_lasso.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./run_lasso.sh -num_dolphin_workers 1 -num_servers 1 -input sample_lasso -features 10 -features_per_partition 5 -lambda 0.5 -step_size 1.0 -max_num_epochs 10 -num_mini_batches 4 -num_worker_blocks 4
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.lasso import main_run as main; sys.exit(main(sys.argv[1:]))" "$@"
