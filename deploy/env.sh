# Environment of the job-server service (sourced by harmony-b200.service).
export PATH=/usr/local/cuda/bin:$PATH
# NVLS (in-switch reductions) only matters for the NCCL baseline arm; the data plane uses its own peer-memory kernels
export NCCL_NVLS_ENABLE=1
# tasklets of different jobs run on their own streams: give them enough hardware queues
export CUDA_DEVICE_MAX_CONNECTIONS=32
# HBM arena segment (bytes) exported over CUDA IPC; tables larger than this allocate additional segments
export HB_ARENA_SEGMENT_BYTES=$((8 << 30))
# torch.distributed rendezvous must not depend on the container hostname
export MASTER_ADDR=127.0.0.1
