#!/bin/bash
# Per-launch metrics of every GEMM launch of one eager f16x3 step (tensor pipe, DRAM and L2 bytes, grid size).
out=gpurun_out
mkdir -p $out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"gemm_tn|gemm_wgrad" --csv --log-file $out/r2b_gemm_launches_f16x3.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --no-backbone --precision f16x3 > /dev/null 2>&1
echo "rc=$?"
wc -l $out/r2b_gemm_launches_f16x3.csv
