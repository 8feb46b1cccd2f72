# Environment of the job-server service: plain KEY=VALUE lines (systemd EnvironmentFile; `set -a; . deploy/env.sh`
# in a shell).
PATH=/usr/local/cuda/bin:/usr/local/bin:/usr/bin:/bin
# NVLS (in-switch reductions) only matters for the NCCL baseline arm; the data plane uses its own peer-memory kernels
NCCL_NVLS_ENABLE=1
# tasklets of different jobs run on their own streams: give them enough hardware queues
CUDA_DEVICE_MAX_CONNECTIONS=32
# HBM arena segment (bytes, 8 GiB) exported over CUDA IPC; larger tables allocate additional segments
HB_ARENA_SEGMENT_BYTES=8589934592
# torch.distributed rendezvous must not depend on the container hostname
MASTER_ADDR=127.0.0.1
