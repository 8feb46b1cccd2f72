#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_gbt.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_gbt.sh -input sample_gbt -metadata_path sample_gbt.meta -features 784 -step_size 0.1 -lambda 0 -gamma 0 -max_depth_of_tree 7 -leaf_min_size 4 -num_keys 5 -max_num_epochs 5 -num_mini_batches 4 -num_worker_blocks 4
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.gbt import main_submit as main; sys.exit(main(sys.argv[1:]))" "$@"
