#!/bin/bash
# harmony_b200 launcher (CLI-compatible with the reference's jobserver/bin/submit_lda.sh; all flags are `-name value`).
# EXAMPLE USAGE
# ./submit_lda.sh -input sample_lda -num_topics 10 -num_vocabs 17935 -max_num_epochs 3 -num_mini_batches 20 -num_worker_blocks 20
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -c "import sys; from harmony_b200.mlapps.lda import main_submit as main; sys.exit(main(sys.argv[1:]))" "$@"
