#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
run() { env "$@" timeout 200 $B $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$EXTRA $*', round(d['value']/1e6,1), 'Mpairs/s frac', round(d['roofline']['frac'],3), 'failed', d['config']['failed_pairs'])"; }
run DCOL_REFILL=0
run DCOL_REFILL=0 DCOL_SIDE_STREAMS=16
for gs in "8 8" "8 16" "8 32" "12 16" "16 16" "16 32" "24 32"; do set -- $gs; run DCOL_REFILL=1 DCOL_REFILL_GEN=$1 DCOL_SIDE_STREAMS=$2; done
EXTRA="--workload config5"
run DCOL_REFILL=0
for gs in "8 16" "16 16" "16 32" "32 32"; do set -- $gs; run DCOL_REFILL=1 DCOL_REFILL_GEN=$1 DCOL_SIDE_STREAMS=$2; done
EXTRA=""
DCOL_REFILL=1 DCOL_REFILL_GEN=8 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 360 -c 120 --csv --log-file gpurun_out/times_refill_v6_g8.csv $B --steps 1 > gpurun_out/times_refill_v6.log 2>&1
