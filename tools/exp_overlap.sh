#!/bin/bash
# 2-GPU experiment: in-graph overlapped gradient all-reduce at full size (hung in round 1). Every run is bounded by
# its own faulthandler exit (70 s) and a SIGTERM timeout; outputs go to files, no pipes.
out=gpurun_out
mkdir -p $out
export PTB200_TRACE_STEPS=1 PTB200_OVERLAP_ALLREDUCE=1 NCCL_DEBUG=WARN
run() {
  tag=$1; shift
  echo "== $tag"
  env "$@" timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $((29500 + RANDOM % 400)) tools/check_ddp.py --height 800 --width 1333 --steps 6 --precision f16 \
    > $out/overlap_$tag.log 2>&1 < /dev/null
  echo "rc=$?"
  grep -v "^W0\|^\*\*\*\|Setting OMP" $out/overlap_$tag.log | tail -${TAILN:-45}
  sleep 2
}
run all148 X=1
TAILN=25 run ctas132 PTB200_GEMM_CTAS=132
