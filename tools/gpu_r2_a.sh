#!/bin/bash
# round-2 GPU call A (2 GPUs): full GPU test suite + multi-GPU worker + short benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/a_gpus.txt 2>&1
timeout 240 python __graft_entry__.py smoke > gpurun_out/a_smoke.txt 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/a_smoke.txt
timeout 420 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/a_multigpu.txt 2>&1; echo "multigpu rc=$?"
tail -25 gpurun_out/a_multigpu.txt
cat gpurun_out/multigpu_report_2gpu.json 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/a_bench2.json 2> gpurun_out/a_bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/a_bench2.err; cat gpurun_out/a_bench2.json
timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/a_bench1.json 2> gpurun_out/a_bench1.err; echo "bench1 rc=$?"
tail -5 gpurun_out/a_bench1.err; cat gpurun_out/a_bench1.json
timeout 700 python -m pytest tests -m gpu -q --deselect tests/test_multigpu.py > gpurun_out/a_pytest.txt 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/a_pytest.txt
