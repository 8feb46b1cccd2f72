#!/bin/bash
# round-2 GPU call L (8 GPUs): headline bench (both modes, NCCL comparators, e2e), LDA with the sparse pull,
# MLR through the job server (process executors, ship mode, leases), 4 concurrent jobs + live re-partition
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/l_clocks.csv 2>/dev/null &
SMI=$!
timeout 300 $TR --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/l_bench8.json 2> gpurun_out/l_bench8.err; echo "bench rc=$?"; tail -2 gpurun_out/l_bench8.err | cut -c1-300; cat gpurun_out/l_bench8.json
timeout 200 $TR --master-port 29712 bench.py --config lda --gpus $N --steps 10 --warmup 3 > gpurun_out/l_lda8.json 2> gpurun_out/l_lda8.err; echo "lda rc=$?"; tail -2 gpurun_out/l_lda8.err | cut -c1-300; cat gpurun_out/l_lda8.json
timeout 200 $TR --master-port 29713 bench.py --config lda --impl nccl --gpus $N --steps 10 --warmup 3 > gpurun_out/l_lda8_nccl.json 2> gpurun_out/l_lda8_nccl.err; echo "lda nccl rc=$?"; cat gpurun_out/l_lda8_nccl.json
timeout 300 python benchmarks/bench_jobserver_mlr.py --executors 8 --epochs 6 --batches_per_worker 8 --executor_mode process > gpurun_out/l_jsmlr8.json 2> gpurun_out/l_jsmlr8.err; echo "jsmlr rc=$?"; tail -3 gpurun_out/l_jsmlr8.err | cut -c1-300; cat gpurun_out/l_jsmlr8.json
timeout 200 python bench.py --config multijob --gpus $N --backend device --scale 1.0 --epochs 6 > gpurun_out/l_multijob8.json 2> gpurun_out/l_multijob8.err; echo "multijob rc=$?"; tail -2 gpurun_out/l_multijob8.err | cut -c1-300; cat gpurun_out/l_multijob8.json
kill $SMI
