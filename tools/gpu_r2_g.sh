#!/bin/bash
# round-2 GPU call G (1 GPU): ncu --set full of the ship-mode kernels under the emulated 8-GPU workload shape
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'ship3_fwd|mlr_bwd_push_warp|ship_softmax' -s 12 -c 6 -o gpurun_out/g_ship_micro python tools/ship_micro.py 8 > gpurun_out/g_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/g_ncu.log
ls -la gpurun_out/g_ship_micro.ncu-rep
