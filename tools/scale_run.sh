#!/bin/bash
# Data-parallel bench exactly as the driver launches it, for N = $1 GPUs (run under `gpurun --gpus N`).
# Keeps the JSON line, the NCCL topology lines and (with LONG=1) a long soak run.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
run() {
  tag=$1; steps=$2
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $N --steps $steps --warmup 5 \
      > $O/r02_scale_${tag}.json 2> $O/r02_scale_${tag}.err
  rc=$?
  echo "== $tag rc=$rc"
  grep -h "NVLS\|Connected all\|via P2P\|comm 0x.*rank 0 nranks" $O/r02_scale_${tag}.err | head -6 | cut -c1-220
  python - "$O/r02_scale_${tag}.json" <<'PY'
import json, sys
try:
    line = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{"metric"')][-1]   # NCCL INFO shares stdout
    d = json.loads(line)
    open(sys.argv[1].replace(".json", ".line.json"), "w").write(line + "\n")
    print(d["n_gpus"], "GPUs:", d["value"], "clips/s", d["ms_per_step"], "ms | e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "ms | dp_check", d.get("dp_check"), "| bytes", d.get("allreduce_bytes_per_step"), d["clocks"])
except Exception as e:
    print("no json:", e)
PY
  grep -h "hangwatch rank\|Traceback\|Error" $O/r02_scale_${tag}.err | head -5
}
run N${N} 20
if [ -n "$LONG" ]; then run N${N}_long ${LONG}; fi
