#!/bin/bash
# 1-GPU battery (round 2, second pass): the unified tf32/fp16 kernels, deterministic wgrad, fused optimizer.
mkdir -p gpurun_out
O=gpurun_out
echo "== CUDA harnesses"
timeout 300 ./build/test_conv_gemm > $O/b_test_conv_gemm.log 2>&1; echo "rc=$?"; tail -3 $O/b_test_conv_gemm.log
timeout 300 ./build/test_wgrad --perf > $O/b_test_wgrad.log 2>&1; echo "rc=$?"; tail -12 $O/b_test_wgrad.log
echo "== pytest: kernels + full size (both operand types)"
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q -s > $O/b_pytest_kernels.log 2>&1; echo "rc=$?"; tail -25 $O/b_pytest_kernels.log
echo "== pytest: modules (tf32)"
timeout 1200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_optim.py tests/test_gpu_benchmark_size.py -m gpu -q -s > $O/b_pytest_modules.log 2>&1; echo "rc=$?"; tail -25 $O/b_pytest_modules.log
echo "== pytest: fp16 path"
timeout 1200 python -m pytest tests/test_gpu_fp16.py -m gpu -q -s > $O/b_pytest_fp16.log 2>&1; echo "rc=$?"; tail -40 $O/b_pytest_fp16.log
echo "== bench tf32 / fp16"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager > $O/b_bench_tf32.json 2> $O/b_bench_tf32.err; echo "rc=$?"
VNQA_OPERANDS=fp16 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager > $O/b_bench_fp16.json 2> $O/b_bench_fp16.err; echo "rc=$?"; tail -3 $O/b_bench_fp16.err
python - <<'PY'
import json
for f in ("b_bench_tf32","b_bench_fp16"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
        for k,v in list(d["kernels"].items())[:12]: print("   ", k, v["ms_per_launch"], v.get("ms_min_med_max"), v.get("tflops", v.get("gb_per_s")))
    except Exception as e: print(f, "no json", e)
PY
echo "battery b done"
