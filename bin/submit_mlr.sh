#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_mlr.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_mlr.sh -input sample_mlr -max_num_epochs 20 -num_mini_batches 10 -num_worker_blocks 10 -step_size 0.1 -classes 10 -features 784 -features_per_partition 392 -model_gaussian 0.001 -lambda 0.005 -decay_period 5 -decay_rate 0.9
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.mlr import main_submit as main; sys.exit(main(sys.argv[1:]))" "$@"
