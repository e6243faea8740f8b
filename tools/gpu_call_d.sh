#!/bin/bash
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
DCOL_REFILL=1 DCOL_REFILL_GEN=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 360 -c 120 --csv --log-file gpurun_out/times_refill_v6_g2.csv $B > gpurun_out/times_refill_v6g2.log 2>&1
DCOL_REFILL=1 DCOL_REFILL_GEN=3 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 360 -c 120 --csv --log-file gpurun_out/times_refill_v6_g3.csv $B > gpurun_out/times_refill_v6g3.log 2>&1
