#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/run_mlr.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./run_mlr.sh -num_dolphin_workers 4 -num_servers 2 -input sample_mlr -max_num_epochs 20 -num_mini_batches 10 -num_worker_blocks 10 -step_size 0.1 -classes 10 -features 784 -features_per_partition 392 -optimizer edu.snu.cay.dolphin.optimizer.impl.EmptyPlanOptimizer
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.mlr import main_run as main; sys.exit(main(sys.argv[1:]))" "$@"
