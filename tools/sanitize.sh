#!/bin/bash
# compute-sanitizer passes over the data-plane kernels (race detection / sanitizers).
# Small shapes only: the tools slow kernels down 10-100x.
mkdir -p gpurun_out
SEL='(pull_push_add and 64) or lda_sparse or sgd_clamp or adagrad or (mlr_step and 784) or migrate_block or device_clock or (nmf_step and 8-0) or ship'
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_apps.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "== $tool: pytest rc=$? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
  tail -2 gpurun_out/sanitizer_${tool}_pytest.log
done
