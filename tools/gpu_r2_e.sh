#!/bin/bash
# round-2 GPU call E (N GPUs): BASELINE configs 3-5 through bench.py --config + v3 ship check
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/e_multigpu_$N.txt 2>&1; echo "multigpu rc=$? t=$(( $(date +%s)-t0 ))"
tail -3 gpurun_out/e_multigpu_$N.txt
timeout 240 $TR --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/e_bench$N.json 2> gpurun_out/e_bench$N.err; echo "bench rc=$? t=$(( $(date +%s)-t0 ))"
tail -2 gpurun_out/e_bench$N.err; cat gpurun_out/e_bench$N.json
for cfg in nmf lda; do
timeout 240 $TR --master-port 29712 bench.py --config $cfg --gpus $N --steps 10 --warmup 3 > gpurun_out/e_${cfg}$N.json 2> gpurun_out/e_${cfg}$N.err; echo "$cfg rc=$? t=$(( $(date +%s)-t0 ))"
tail -2 gpurun_out/e_${cfg}$N.err; cat gpurun_out/e_${cfg}$N.json
timeout 240 $TR --master-port 29713 bench.py --config $cfg --impl nccl --gpus $N --steps 10 --warmup 3 > gpurun_out/e_${cfg}${N}_nccl.json 2> gpurun_out/e_${cfg}${N}_nccl.err; echo "$cfg nccl rc=$? t=$(( $(date +%s)-t0 ))"
tail -2 gpurun_out/e_${cfg}${N}_nccl.err; cat gpurun_out/e_${cfg}${N}_nccl.json
done
timeout 300 python bench.py --config multijob --gpus $N --backend device --scale 1.0 --epochs 6 > gpurun_out/e_multijob$N.json 2> gpurun_out/e_multijob$N.err; echo "multijob rc=$? t=$(( $(date +%s)-t0 ))"
tail -4 gpurun_out/e_multijob$N.err; cat gpurun_out/e_multijob$N.json
