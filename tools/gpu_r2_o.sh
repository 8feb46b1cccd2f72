#!/bin/bash
# round-2 GPU call O (1 GPU): the full GPU suite, twice, in fresh processes (hunting a GC-time segfault seen once)
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python -X faulthandler -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/o_pytest$i.txt 2>&1; echo "pytest run $i rc=$?"
tail -2 gpurun_out/o_pytest$i.txt | cut -c1-200; grep -n "^FAILED\|^ERROR\|Segmentation\|Garbage" gpurun_out/o_pytest$i.txt | head -5
done
