#!/bin/bash
# compute-sanitizer over the K12 / K9 kernels and the rewritten LDA sampler (small shapes).
mkdir -p gpurun_out
SEL='lasso_cd_kernel or gbt_device_tree or lgamma_sum or lda_sampler'
for tool in memcheck racecheck; do
  timeout 250 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_k12_$tool.log \
      python -m pytest tests/test_gpu_apps.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_k12_${tool}_pytest.log 2>&1
  echo "== $tool: pytest rc=$? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_k12_$tool.log | tail -1)"
  tail -1 gpurun_out/sanitizer_k12_${tool}_pytest.log
done
