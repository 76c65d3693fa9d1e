#!/bin/bash
# 1-GPU battery (round 2, third pass): wgrad with the fill-based split count, both operand types end to end.
mkdir -p gpurun_out
O=gpurun_out
echo "== wgrad harness"
timeout 300 ./build/test_wgrad --perf > $O/c_test_wgrad.log 2>&1; echo "rc=$?"; grep "P wgrad\|ALL\|FAIL" $O/c_test_wgrad.log
echo "== pytest"
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_modules.py tests/test_gpu_optim.py tests/test_gpu_benchmark_size.py tests/test_gpu_fp16.py -m gpu -q -s > $O/c_pytest.log 2>&1; echo "rc=$?"; tail -12 $O/c_pytest.log
grep -h "^wgrad act\|fp16:\|fp16 logits\|fp16 B=64\|logits: vs\|B=64 train\|reference under" $O/c_pytest.log | cut -c1-200
echo "== bench tf32 / fp16"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --operands tf32 > $O/c_bench_tf32.json 2> $O/c_bench_tf32.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --operands fp16 > $O/c_bench_fp16.json 2> $O/c_bench_fp16.err; echo "rc=$?"; tail -3 $O/c_bench_fp16.err
python - <<'PY'
import json
for f in ("c_bench_tf32","c_bench_fp16"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], d.get("gpu_eager_baseline"))
        for k,v in list(d["kernels"].items())[:14]: print("   ", k, v["ms_per_launch"], v.get("ms_min_med_max"), v.get("tflops", v.get("gb_per_s")), v["share_of_step"])
    except Exception as e: print(f, "no json", e)
PY
echo "== argmax gate, fp16 operands"
timeout 900 python tests/tools/argmax_agreement.py --clips 4096 --operands fp16 > $O/c_argmax_fp16.log 2>&1; echo "rc=$?"; tail -3 $O/c_argmax_fp16.log
echo "battery c done"
