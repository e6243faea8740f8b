#!/bin/bash
# per-kernel A/B: 4 CTAs/SM (255 registers) against 6 CTAs/SM (168 registers), each of the 40 config-4 launches alone
L="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample --no-python-reference"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pair_kernel -s 120 -c 40 --csv --log-file gpurun_out/launches_mb4.csv $L > /dev/null 2>&1
DCOL_LIB=$PWD/dcol_trajectory_optimization_b200/libdcol_b200_mb6.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pair_kernel -s 120 -c 40 --csv --log-file gpurun_out/launches_mb6.csv $L > /dev/null 2>&1
DCOL_LIB=$PWD/dcol_trajectory_optimization_b200/libdcol_b200_mb6.so $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mb6 whole step', round(d['value']/1e6,1), d['ms_per_step'])"
python - <<'PY'
import csv
def load(f):
    out={}
    for r in csv.reader(open(f)):
        if len(r)>10 and r[0].isdigit(): out[r[4]]=float(r[-1])
    return out
a,b=load("gpurun_out/launches_mb4.csv"),load("gpurun_out/launches_mb6.csv")
for k in a:
    if k in b: print(f"{k[17:60]:45s} {a[k]/1e3:8.1f} {b[k]/1e3:8.1f} us  ratio {b[k]/a[k]:.3f}")
print("sum", sum(a.values())/1e6, sum(b[k] for k in a if k in b)/1e6)
PY
