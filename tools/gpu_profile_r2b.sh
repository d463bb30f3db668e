#!/bin/bash
# Second profile pass of round 2 (one gpurun call, final build): launch lists of the f16x3 and f16 steps and the
# per-launch tensor-pipe / DRAM / L2 metrics of every GEMM launch of one eager f16x3 step.
out=gpurun_out
mkdir -p $out
for prec in f16x3 f16; do
  echo "== launch list $prec"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2b_launches_$prec.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --no-backbone --precision $prec > /dev/null 2>&1
  echo "rc=$?"; python tools/agg_launches.py $out/r2b_launches_$prec.csv 13 > $out/r2b_launches_step_$prec.txt 2>&1; head -16 $out/r2b_launches_step_$prec.txt
done
bash tools/gpu_profile_gemm_launches.sh
