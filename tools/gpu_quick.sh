#!/bin/bash
# quick GPU check of a kernel change: GPU tests (without the 90 s quadrotor scenario), then the device-resident rate of
# config 4 (with the parity sample against the oracle) and config 5.  usage: bash tools/gpu_quick.sh [tag]
TAG=${1:-quick}
mkdir -p gpurun_out
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-python-reference"
DCOL_SKIP_SLOW=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py $Q > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 300 python bench.py --workload config5 $Q --no-parity-sample > gpurun_out/bench_${TAG}_c5.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("config4", round(d["value"]/1e6,1), "M pairs/s", round(d["ms_per_step"],3), "ms  frac", round(d["roofline"]["frac"],4))
p=d.get("parity_sample") or {}
print("parity", {k:p.get(k) for k in ("pairs","status_mismatches","iteration_count_mismatches","max_alpha_rel_err","pairs_with_grad_err_above_1e-6","rounding_sensitive_pairs","max_grad_rel_err")})
d=json.loads(open("gpurun_out/bench_${TAG}_c5.json").read().strip().splitlines()[-1])
print("config5", round(d["value"]/1e6,1), "M pairs/s", round(d["ms_per_step"],3), "ms  frac", round(d["roofline"]["frac"],4))
PY
