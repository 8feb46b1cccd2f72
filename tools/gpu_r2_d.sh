#!/bin/bash
# round-2 GPU call D (1 GPU): ship kernels under the emulated 8-GPU workload shape + 1-GPU tests of the new paths
mkdir -p gpurun_out
timeout 200 python tools/ship_micro.py 8 > gpurun_out/d_micro8.json 2> gpurun_out/d_micro8.err; echo "micro rc=$?"; tail -3 gpurun_out/d_micro8.err; cat gpurun_out/d_micro8.json
HB_SHIP_K1=cta timeout 200 python tools/ship_micro.py 8 > gpurun_out/d_micro8_cta.json 2> gpurun_out/d_micro8_cta.err; cat gpurun_out/d_micro8_cta.json
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "ship or lazy or migration or guard" 2>&1 | tail -5
