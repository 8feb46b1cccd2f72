#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_nmf.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_nmf.sh -input sample_nmf -rank 30 -features 100 -step_size 0.01 -max_num_epochs 50 -num_mini_batches 25 -num_worker_blocks 25 -decay_rate 0.9 -decay_period 5
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.nmf import main_submit as main; sys.exit(main(sys.argv[1:]))" "$@"
