#!/bin/bash
# round-2 GPU call J (1 GPU): full GPU suite + LDA (sparse pull) bench + smoke
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_multigpu.py > gpurun_out/j_pytest.txt 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/j_pytest.txt | cut -c1-300
grep -n "^FAILED\|^ERROR" gpurun_out/j_pytest.txt | head
timeout 200 python bench.py --config lda --gpus 1 --steps 10 --warmup 3 > gpurun_out/j_lda1.json 2> gpurun_out/j_lda1.err; echo "lda rc=$?"; tail -2 gpurun_out/j_lda1.err; cat gpurun_out/j_lda1.json
timeout 200 python bench.py --config nmf --gpus 1 --steps 10 --warmup 3 > gpurun_out/j_nmf1.json 2> gpurun_out/j_nmf1.err; echo "nmf rc=$?"; tail -2 gpurun_out/j_nmf1.err; cat gpurun_out/j_nmf1.json
