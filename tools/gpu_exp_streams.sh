#!/bin/bash
# A/B: hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) x side streams on the host-entry chunk sweep
mkdir -p gpurun_out
for C in 8 32; do for S in 8 16 32; do
  export CUDA_DEVICE_MAX_CONNECTIONS=$C DCOL_SIDE_STREAMS=$S
  timeout 300 python tools/diag_e2e.py 2>&1 | grep "side streams" | sed "s/^/conn $C /"
done; done
