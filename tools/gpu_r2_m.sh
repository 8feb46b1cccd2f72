#!/bin/bash
# round-2 GPU call M (8 GPUs): job-server MLR and the 4-job scenario after the fence / rendezvous fixes
mkdir -p gpurun_out
timeout 300 python benchmarks/bench_jobserver_mlr.py --executors 8 --epochs 10 --batches_per_worker 8 --executor_mode process --spmd_value 9112.5 > gpurun_out/m_jsmlr8.json 2> gpurun_out/m_jsmlr8.err; echo "jsmlr rc=$?"; tail -3 gpurun_out/m_jsmlr8.err | cut -c1-300; cat gpurun_out/m_jsmlr8.json
timeout 200 python bench.py --config multijob --gpus 8 --backend device --scale 1.0 --epochs 6 > gpurun_out/m_multijob8.json 2> gpurun_out/m_multijob8.err; echo "multijob rc=$?"; tail -2 gpurun_out/m_multijob8.err | cut -c1-300; cat gpurun_out/m_multijob8.json
