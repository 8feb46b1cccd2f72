#!/bin/bash
# round-2 GPU call N (1 GPU): final full GPU suite + smoke + dense bench (both tcgen05 variants with clocks)
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.txt 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/n_pytest.txt | cut -c1-300; grep -n "^FAILED\|^ERROR" gpurun_out/n_pytest.txt | head
timeout 200 python bench.py --config dense --steps 20 --warmup 5 > gpurun_out/n_dense.json 2> gpurun_out/n_dense.err; echo "dense rc=$?"; cat gpurun_out/n_dense.json | cut -c1-1500
