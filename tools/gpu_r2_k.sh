#!/bin/bash
# round-2 GPU call K (2 GPUs): elasticity-on-device tests, 2-rank data plane, LDA sparse pull vs NCCL arm,
# job-server MLR vs SPMD, ncu NVLink byte counters of the fused kernels (ONE process driving two GPUs)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_jobserver.py tests/test_gpu_kernels.py -m gpu -q -k "reconfigure_workers_live_on_device or lda_sparse or sm_share" 2>&1 | tail -4 | cut -c1-300
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -2
for impl in harmony nccl; do
timeout 200 $TR --master-port 29712 bench.py --config lda --impl $impl --gpus 2 --steps 10 --warmup 3 > gpurun_out/k_lda2_$impl.json 2> gpurun_out/k_lda2_$impl.err; echo "lda $impl rc=$?"; tail -2 gpurun_out/k_lda2_$impl.err; cut -c1-420 gpurun_out/k_lda2_$impl.json
done
timeout 300 python benchmarks/bench_jobserver_mlr.py --executors 2 --epochs 6 --batches_per_worker 4 > gpurun_out/k_jsmlr2.json 2> gpurun_out/k_jsmlr2.err; echo "jsmlr rc=$?"; tail -3 gpurun_out/k_jsmlr2.err | cut -c1-300; cat gpurun_out/k_jsmlr2.json
timeout 300 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'mlr_fwd|mlr_bwd|migrate|flip' --csv --log-file gpurun_out/k_nvlink.csv python tools/nvlink_probe.py > gpurun_out/k_nvlink_probe.log 2>&1; echo "ncu nvlink rc=$?"; tail -2 gpurun_out/k_nvlink_probe.log | cut -c1-300
grep -c "nvl" gpurun_out/k_nvlink.csv; head -30 gpurun_out/k_nvlink.csv | cut -c1-260
