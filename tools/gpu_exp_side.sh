#!/bin/bash
# device-resident rate against the number of side streams (how many different pair kernels share the SMs)
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-python-reference --no-parity-sample"
for S in 1 2 3 4 6 8 12; do
  for W in config4 config5; do
    v=$(DCOL_SIDE_STREAMS=$S timeout 300 python bench.py --workload $W $Q 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']/1e6,1), round(d['ms_per_step'],3))")
    echo "side $S $W: $v"
  done
done
