#!/bin/bash
# 1-GPU battery (round 2, fifth pass): the GPU suite exactly as the driver runs it, the default bench line, and the
# re-tiled streaming kernels under ncu.
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu (driver command)"
timeout 2700 python -m pytest tests/ -x -q -m gpu > $O/e_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/e_pytest_gpu.log
echo "== bench (defaults)"
timeout 900 python bench.py > $O/e_bench_default.json 2> $O/e_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/e_bench_default.json")); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline_hbm"], d["clocks"])
PY
echo "== ncu: streaming kernels"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'film_bwd|bn_apply|frames_to|bn_bwd' \
    --csv --log-file $O/r02_streaming_kernels.csv python tools/ncu_targets.py --operands fp16 > $O/e_ncu_stream.log 2>&1; echo "rc=$?"
grep -v "^==" $O/r02_streaming_kernels.csv | cut -d, -f5,13- | head -20
echo "battery e done"
