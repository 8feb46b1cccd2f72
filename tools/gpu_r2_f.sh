#!/bin/bash
# round-2 GPU call F (1 GPU): stream-K tcgen05 numerics + TFLOP/s, ship micro with the tile-per-CTA K1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "dense or ship or lazy" 2>&1 | tail -8
timeout 200 python tools/ship_micro.py 8 > gpurun_out/f_micro8.json 2> gpurun_out/f_micro8.err; echo "micro rc=$?"; tail -3 gpurun_out/f_micro8.err; cat gpurun_out/f_micro8.json
timeout 300 python bench.py --config dense --steps 20 --warmup 5 > gpurun_out/f_dense.json 2> gpurun_out/f_dense.err; echo "dense rc=$?"; tail -3 gpurun_out/f_dense.err; cat gpurun_out/f_dense.json
