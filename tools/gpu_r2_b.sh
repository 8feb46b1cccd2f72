#!/bin/bash
# round-2 GPU call B (N GPUs): multi-GPU worker + bench at N (+ optional N=1)
N=${1:-2}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/b_multigpu_$N.txt 2>&1; echo "multigpu rc=$?"
tail -5 gpurun_out/b_multigpu_$N.txt
cat gpurun_out/multigpu_report_${N}gpu.json 2>/dev/null; echo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/b_bench$N.json 2> gpurun_out/b_bench$N.err; echo "bench$N rc=$?"
tail -5 gpurun_out/b_bench$N.err; cat gpurun_out/b_bench$N.json
if [ -n "$2" ]; then
timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b_bench1.json 2> gpurun_out/b_bench1.err; echo "bench1 rc=$?"
tail -5 gpurun_out/b_bench1.err; cat gpurun_out/b_bench1.json
fi
