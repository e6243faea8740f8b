#!/bin/bash
mkdir -p gpurun_out
python tools/debug_grad.py 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q -x -k "lane_refill or random_shapes or face_count or scene or extensions or jacobian_vs_dense" > gpurun_out/pytest_r2_c.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2_c.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
for v in "DCOL_REFILL=0" "DCOL_REFILL=1 DCOL_REFILL_GEN=1" "DCOL_REFILL=1 DCOL_REFILL_GEN=2" "DCOL_REFILL=1 DCOL_REFILL_GEN=4" "DCOL_REFILL=1 DCOL_REFILL_GEN=8" "DCOL_REFILL=1 DCOL_REFILL_GEN=16" "DCOL_REFILL=1 DCOL_REFILL_GEN=32"; do
  env $v timeout 200 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']/1e6,1), 'Mpairs/s frac', round(d['roofline']['frac'],3), 'failed', d['config']['failed_pairs'])"
done
for v in "DCOL_REFILL=0" "DCOL_REFILL=1 DCOL_REFILL_GEN=4" "DCOL_REFILL=1 DCOL_REFILL_GEN=16"; do
  env $v timeout 200 $B --workload config5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $v', round(d['value']/1e6,1), 'Mpairs/s frac', round(d['roofline']['frac'],3), 'failed', d['config']['failed_pairs'])"
done
DCOL_REFILL=1 DCOL_REFILL_GEN=8 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -s 360 -c 120 --csv --log-file gpurun_out/times_refill_v5_g8.csv $B --steps 1 > gpurun_out/times_refill_v5.log 2>&1
DCOL_REFILL=1 DCOL_REFILL_GEN=8 ncu --set full --clock-control none --import-source on -k regex:trip_kernel -s 120 -c 2 -o gpurun_out/prof_r2_refill_v5 $B --steps 1 > gpurun_out/ncu_r2_v5.log 2>&1
ls -la gpurun_out/prof_r2_refill_v5.ncu-rep
