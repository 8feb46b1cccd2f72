#!/bin/bash
# round-2 GPU call H (2 GPUs): micro (1 GPU) + 2-rank correctness + 2-GPU bench
mkdir -p gpurun_out
timeout 200 python tools/ship_micro.py 8 > gpurun_out/h_micro8.json 2> gpurun_out/h_micro8.err; echo "micro rc=$?"; tail -3 gpurun_out/h_micro8.err; cat gpurun_out/h_micro8.json
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/h_multigpu_2.txt 2>&1; echo "multigpu rc=$?"
tail -3 gpurun_out/h_multigpu_2.txt; grep -n "Error:" gpurun_out/h_multigpu_2.txt | head -5
cat gpurun_out/multigpu_report_2gpu.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 --no_nccl > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/h_bench2.err; cat gpurun_out/h_bench2.json
