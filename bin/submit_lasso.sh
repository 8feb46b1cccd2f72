#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_lasso.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_lasso.sh -input sample_lasso -features 10 -features_per_partition 5 -lambda 0.5 -step_size 1.0 -max_num_epochs 10 -num_mini_batches 4 -num_worker_blocks 4
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.lasso import main_submit as main; sys.exit(main(sys.argv[1:]))" "$@"
