#!/bin/bash
# A/B of the lane-refill kernels on one GPU (writes gpurun_out/ab2_*.log).  Usage: bash tools/ab_refill.sh
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
run() { name=$1; shift; echo "== $name"; ( timeout 180 env "$@" $B $EXTRA_ARGS > gpurun_out/ab2_$name.log 2> gpurun_out/ab2_$name.err; echo "rc=$?" ) ; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab2_$name.log").read().strip().splitlines()[-1])
    print("$name", round(d["value"]/1e6,1), "Mpairs/s", "frac", round(d["roofline"]["frac"],3), "failed", d["config"]["failed_pairs"])
except Exception as e:
    print("$name", "no result", e)
PY
}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/ab2_pytest.log
run plain DCOL_REFILL=0
run refill_auto DCOL_REFILL=1
for g in 1 2 4 8 16; do run refill_g$g DCOL_REFILL=1 DCOL_REFILL_GEN=$g; done
EXTRA_ARGS="--workload config5"
run c5_plain DCOL_REFILL=0
run c5_refill_auto DCOL_REFILL=1
run c5_refill_g4 DCOL_REFILL=1 DCOL_REFILL_GEN=4
run c5_refill_g16 DCOL_REFILL=1 DCOL_REFILL_GEN=16
