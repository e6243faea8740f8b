#!/bin/bash
# one-GPU validation + measurement pass: GPU tests, the default bench line, config 5, the reference arm, ncu launch list and
# a full ncu capture of the first specialisations of the timed step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2_final.log
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_n1.err
timeout 600 python bench.py --workload config5 --no-altro --no-cpu-baseline --no-jacobian --no-coherent --no-sizes --no-parity-sample > gpurun_out/bench_r2_n1_config5.json 2>/dev/null; echo "bench c5 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; echo "bench ref rc=$?"
L="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-altro --no-jacobian --no-coherent --no-sizes --no-parity-sample"
ncu --metrics gpu__time_duration.sum --clock-control none -s 125 -c 45 --csv --log-file gpurun_out/r02_launches_ncu_time_duration.csv $L > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 120 -c 14 -o gpurun_out/prof_r2_final $L > gpurun_out/ncu_full_r2.log 2>&1
ls -la gpurun_out/prof_r2_final.ncu-rep gpurun_out/r02_launches_ncu_time_duration.csv
python - <<'PY'
import json
for f in ("bench_r2_n1","bench_r2_n1_config5","bench_r2_reference"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2), "M/s", "frac", (d.get("roofline") or {}).get("frac"), "e2e", (d.get("e2e") or {}).get("value"), "scene", (d.get("e2e_scene") or {}).get("value"), (d.get("e2e_scene") or {}).get("e2e_over_device"))
        for k in ("parity_sample","cpu_baseline_python","parity_vs_python_reference","config5_full","config4_2p26","altro"):
            if k in d: print("   ",k, json.dumps(d[k])[:400])
    except Exception as e: print(f,"no result",e)
PY
