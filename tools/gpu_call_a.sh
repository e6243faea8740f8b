#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_b.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2_b.log
timeout 600 bash tools/run_n.sh 2 > gpurun_out/bench_r2_n2.log 2> gpurun_out/bench_r2_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_r2_n2.log").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gather_verified","slots_checked")}, d.get("config5_full"), d.get("config4_2p26"))
except Exception as e: print("no n2 result", e)
PY
tail -3 gpurun_out/bench_r2_n2.err
export CUDA_VISIBLE_DEVICES=0
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
for v in "DCOL_REFILL=0" "DCOL_REFILL=1 DCOL_REFILL_GEN=1" "DCOL_REFILL=1 DCOL_REFILL_GEN=4"; do
  env $v timeout 200 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']/1e6,1), 'Mpairs/s frac', round(d['roofline']['frac'],3))"
done
