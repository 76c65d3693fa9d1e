#!/bin/bash
# 1-GPU battery (round 2, first pass): A/B of the TMEM-allocation fix under SM sharing, the full GPU test suite, a long
# bench run with the stall monitor armed, the argmax gate, the PyTorch-eager baseline and the fp16 candidate model check.
mkdir -p gpurun_out
O=gpurun_out
echo "== stress, PRE-FIX library (expected: HANG)"; 
LD_LIBRARY_PATH=$PWD/build/prefix timeout 120 ./build/stress_cluster_share --iters 3000 > $O/stress_prefix.log 2>&1; echo "rc=$?"; tail -2 $O/stress_prefix.log
LD_LIBRARY_PATH=$PWD/build/prefix timeout 120 ./build/stress_cluster_share --iters 6000 --n-img 560 > $O/stress_prefix_small.log 2>&1; echo "rc=$?"; tail -2 $O/stress_prefix_small.log
echo "== stress, fixed library (expected: ALL OK)"
timeout 300 ./build/stress_cluster_share --iters 5000 > $O/stress_fixed.log 2>&1; echo "rc=$?"; tail -2 $O/stress_fixed.log
timeout 300 ./build/stress_cluster_share --iters 12000 --n-img 560 > $O/stress_fixed_small.log 2>&1; echo "rc=$?"; tail -2 $O/stress_fixed_small.log
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x -s > $O/pytest_gpu_a.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_gpu_a.log
echo "== pytest -m gpu (continue past first failure, new full-size files only)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_benchmark_size.py -m gpu -q -s > $O/pytest_gpu_fullsize.log 2>&1; echo "rc=$?"; tail -30 $O/pytest_gpu_fullsize.log
echo "== long bench, stall monitor armed"
BENCH_HANGWATCH_S=30 BENCH_VERBOSE=1 timeout 600 python bench.py --steps 200 --warmup 5 > $O/bench_long_a.json 2> $O/bench_long_a.err; echo "rc=$?"; tail -c 600 $O/bench_long_a.json
echo "== argmax agreement gate"
timeout 900 python tests/tools/argmax_agreement.py --clips 4096 > $O/argmax_agreement.log 2>&1; echo "rc=$?"; tail -3 $O/argmax_agreement.log
echo "== eager baseline"
timeout 600 python tests/tools/eager_gpu_baseline.py --batch 64 --steps 3 > $O/eager_baseline.log 2>&1; echo "rc=$?"; tail -2 $O/eager_baseline.log
echo "== fp16 candidate model check"
timeout 600 python tests/tools/check_model_f16.py --full --bench > $O/check_model_f16.log 2>&1; echo "rc=$?"; tail -12 $O/check_model_f16.log
echo "battery done"
