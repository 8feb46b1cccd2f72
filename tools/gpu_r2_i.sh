#!/bin/bash
# round-2 GPU call I (8 GPUs): 8-rank correctness + ship variants on the same box
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/i_multigpu_8.txt 2>&1; echo "multigpu rc=$?"
tail -2 gpurun_out/i_multigpu_8.txt; grep -n "Error:" gpurun_out/i_multigpu_8.txt | head -3; cat gpurun_out/multigpu_report_8gpu.json; echo
run() { name=$1; shift; timeout 200 env "$@" $TR --master-port 29712 bench.py --gpus $N --steps 20 --warmup 5 --mode ship --no_nccl --no_e2e $EXTRA > gpurun_out/i_$name.json 2> gpurun_out/i_$name.err; echo "$name rc=$?"; cat gpurun_out/i_$name.json; }
EXTRA="" run v3 HB_X=1
EXTRA="" run v3_full HB_SHIP_K1_SPARE=0
EXTRA="--ship_impl v1" run v1 HB_X=1
EXTRA="--ship_rs nvls" run v2nvls HB_X=1
