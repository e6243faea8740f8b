#!/bin/bash
# bench at N GPUs of this box, as the driver launches it.  Usage: bash tools/run_n.sh N [extra bench args]
N=$1; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@"
