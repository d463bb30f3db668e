#!/bin/bash
# 2-GPU (or N-GPU: NGPU=8) experiment: in-graph overlapped gradient all-reduce against the single eager all-reduce.
# Every run is bounded by bench.py's own watchdog and a SIGTERM timeout; outputs go to files, no pipes.
out=gpurun_out
n=${NGPU:-2}
mkdir -p $out
export PTB200_WATCHDOG_S=${WATCHDOG:-150}
run() {
  tag=$1; shift
  env "$@" timeout -k 5 $((PTB200_WATCHDOG_S + 30)) python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
    --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps ${STEPS:-60} --warmup 3 \
    --no-cpu-baseline --no-backbone > $out/overlap_${tag}_${n}gpu.json 2> $out/overlap_${tag}_${n}gpu.err < /dev/null
  echo "== $tag rc=$?"
  python - "$out/overlap_${tag}_${n}gpu.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("f16x3 ms/step", round(d["ms_per_step"], 3), "it/s", round(d["value"], 2), "| f16 ms/step",
          round(d["mixed_precision_f16"]["ms_per_step"], 3), "it/s", round(d["mixed_precision_f16"]["value"], 2))
except Exception as e:
    print("no line:", e)
PY
  grep -n "Timeout\|Error\|error" $out/overlap_${tag}_${n}gpu.err | head -5
}
run overlap PTB200_OVERLAP_ALLREDUCE=1
[ -n "$ONLY_OVERLAP" ] && exit 0
run eager PTB200_OVERLAP_ALLREDUCE=0
[ "$n" = "2" ] && run overlap2 PTB200_OVERLAP_ALLREDUCE=1
