#!/bin/bash
# 1-GPU battery (round 2, fourth pass): full GPU suite as the driver runs it, A/B of the generator-backward placement,
# the other configs (eval forward, gp, multi-hop), ncu launch list + full captures, compute-sanitizer.
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu (driver command)"
timeout 2400 python -m pytest tests/ -x -q -m gpu > $O/d_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 $O/d_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/d_smoke.log
echo "== bench fp16: generator backward in line (default) vs side stream"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/d_bench_fp16.json 2> $O/d_bench_fp16.err; echo "rc=$?"
VNQA_SIDE_BWD=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager > $O/d_bench_fp16_sidebwd.json 2> $O/d_bench_fp16_sidebwd.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --operands tf32 > $O/d_bench_tf32.json 2> $O/d_bench_tf32.err; echo "rc=$?"
echo "== other configs: eval forward (config 2), gp, multi-hop, ragged"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload eval > $O/d_bench_eval.json 2> $O/d_bench_eval.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --model gp > $O/d_bench_gp.json 2> $O/d_bench_gp.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --model tmh > $O/d_bench_tmh.json 2> $O/d_bench_tmh.err; echo "rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --ragged > $O/d_bench_ragged.json 2> $O/d_bench_ragged.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("d_bench_fp16","d_bench_fp16_sidebwd","d_bench_tf32","d_bench_eval","d_bench_gp","d_bench_tmh","d_bench_ragged"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "tflops", d["model_tflops"], d["clocks"])
        if f=="d_bench_fp16":
            for k,v in list(d["kernels"].items())[:14]: print("   ", k, v["ms_per_launch"], v.get("ms_min_med_max"), v.get("tflops", v.get("gb_per_s")), v["share_of_step"])
            print("   cpu", d.get("cpu_baseline"), "eager", d.get("gpu_eager_baseline"))
    except Exception as e: print(f, "no json", e)
PY
echo "== ncu launch list of a bench run"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_bench_step.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-eager > $O/d_bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu --set full of the hot kernels (fp16)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm_kernel|wgrad_kernel|film_bwd|bn_apply|frames_to' \
    -o $O/r02_prof_fp16 python tools/ncu_targets.py --operands fp16 > $O/d_ncu_full.log 2>&1; echo "rc=$?"; tail -2 $O/d_ncu_full.log
echo "== compute-sanitizer"
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool ./build/test_wgrad > $O/r02_sanitizer_${tool}_wgrad.log 2>&1; echo "$tool wgrad rc=$?"; tail -3 $O/r02_sanitizer_${tool}_wgrad.log
  timeout 600 compute-sanitizer --tool $tool ./build/test_conv_gemm > $O/r02_sanitizer_${tool}_conv.log 2>&1; echo "$tool conv rc=$?"; tail -3 $O/r02_sanitizer_${tool}_conv.log
done
timeout 300 compute-sanitizer --tool racecheck ./build/test_wgrad --case wgrad3x3_small > $O/r02_sanitizer_racecheck_wgrad.log 2>&1; echo "racecheck wgrad rc=$?"; tail -3 $O/r02_sanitizer_racecheck_wgrad.log
echo "battery d done"
