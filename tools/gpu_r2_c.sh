#!/bin/bash
# round-2 GPU call C (8 GPUs): 8-rank correctness + headline bench + A/B variants of the ship step
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/c_multigpu_$N.txt 2>&1; echo "multigpu rc=$?"
tail -5 gpurun_out/c_multigpu_$N.txt
cat gpurun_out/multigpu_report_${N}gpu.json 2>/dev/null; echo
timeout 400 $TR --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c_bench$N.json 2> gpurun_out/c_bench$N.err; echo "bench$N rc=$?"
tail -3 gpurun_out/c_bench$N.err; cat gpurun_out/c_bench$N.json
for v in cta warp; do
HB_SHIP_K1=$v timeout 300 $TR --master-port 29712 bench.py --gpus $N --steps 20 --warmup 5 --mode ship --no_nccl --no_e2e > gpurun_out/c_bench${N}_k1$v.json 2> gpurun_out/c_bench${N}_k1$v.err; echo "k1=$v rc=$?"
cat gpurun_out/c_bench${N}_k1$v.json
done
timeout 300 $TR --master-port 29713 bench.py --gpus $N --steps 20 --warmup 5 --mode ship --no_nccl --no_e2e --no_nvls > gpurun_out/c_bench${N}_unicast.json 2> gpurun_out/c_bench${N}_unicast.err; echo "unicast rc=$?"
cat gpurun_out/c_bench${N}_unicast.json
timeout 300 $TR --master-port 29714 bench.py --gpus $N --steps 20 --warmup 5 --mode ship --no_nccl --no_e2e --ship_impl v1 > gpurun_out/c_bench${N}_v1.json 2> gpurun_out/c_bench${N}_v1.err; echo "v1 rc=$?"
cat gpurun_out/c_bench${N}_v1.json
