#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu (all, no -x)"
timeout 2700 python -m pytest tests/ -q -m gpu > $O/f_pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/f_pytest_gpu.log
echo "== bench (defaults)"
timeout 900 python bench.py > $O/f_bench_default.json 2> $O/f_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/f_bench_default.json")); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline_hbm"], d["clocks"])
for k,v in list(d["kernels"].items())[:14]: print("   ", k, v["ms_per_launch"], v.get("ms_min_med_max"), v.get("tflops", v.get("gb_per_s")), v["share_of_step"])
PY
echo "battery f done"
