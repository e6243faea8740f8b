#!/bin/bash
# usage: bash tools/gpu_call_n.sh N   -> gpurun_out/bench_r2_n$N.json (+ the 2-GPU gather tests when N >= 2)
N=$1
mkdir -p gpurun_out
timeout 900 bash tools/run_n.sh $N > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r2_n$N.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gather_verified","slots_checked","n_gpus")})
    print("e2e", d.get("e2e")); print("c5", d.get("config5_full")); print("2p26", d.get("config4_2p26")); print(d["config"]["collective"][:120])
except Exception as e: print("no result", e)
PY
tail -2 gpurun_out/bench_r2_n$N.err
