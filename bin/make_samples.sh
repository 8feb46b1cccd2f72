#!/bin/bash
# Generates the sample inputs the EXAMPLE USAGE lines reference (sample_mlr, sample_nmf, sample_lda, sample_lasso,
# sample_gbt(.meta), *_test, sample_host_to_bandwidth, inputs/adj_list, inputs/shortest_path) into this directory.
DIR="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
export PYTHONPATH="$DIR:$PYTHONPATH"
export HB_BIN_DIR="$DIR/bin"
exec python -m harmony_b200.common.samples "${1:-$DIR/bin}"
