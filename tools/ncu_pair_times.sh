#!/bin/bash
# per-launch durations of the 40 pair kernels of one timed step, plain vs lane-refill (serialised by ncu)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
DCOL_REFILL=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pair_kernel -s 120 -c 40 --csv --log-file gpurun_out/times_plain.csv $B > gpurun_out/times_plain.log 2>&1
DCOL_REFILL=1 DCOL_REFILL_GEN=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pair_kernel -s 120 -c 40 --csv --log-file gpurun_out/times_refill_g1.csv $B > gpurun_out/times_refill_g1.log 2>&1
DCOL_REFILL=1 DCOL_REFILL_GEN=4 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pair_kernel -s 120 -c 40 --csv --log-file gpurun_out/times_refill_g4.csv $B > gpurun_out/times_refill_g4.log 2>&1
DCOL_REFILL=1 DCOL_REFILL_GEN=1 ncu --set full --clock-control none --import-source on -k regex:pair_kernel_refill -s 120 -c 2 -o gpurun_out/prof_r2_refill_g1 $B > gpurun_out/ncu_r2_g1.log 2>&1
ls -la gpurun_out/times_*.csv gpurun_out/prof_r2_refill_g1.ncu-rep
