#!/bin/bash
# Round-2 profile pass (one gpurun call): launch list of the f16x3 and f16 steps, tensor-pipe / DRAM metrics per kernel
# of one f16x3 step, `ncu --set full` of the dominant kernel, and the tensor-pipe share of the backbone micro-benchmark.
out=gpurun_out
mkdir -p $out
py=python
for prec in f16x3 f16; do
  echo "== launch list $prec"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2_launches_$prec.csv \
    $py bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --no-backbone --precision $prec > /dev/null 2>&1
  echo "rc=$?"; $py tools/agg_launches.py $out/r2_launches_$prec.csv 13 > $out/r2_launches_step_$prec.txt 2>&1; head -14 $out/r2_launches_step_$prec.txt
done
echo "== per-kernel tensor pipe + dram, f16x3 (one eager step)"
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"gemm_tn|gemm_wgrad|conv1_u8" --csv --log-file $out/r2_gemm_metrics_f16x3.csv \
  $py bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --no-backbone --precision f16x3 > /dev/null 2>&1
echo "rc=$?"
echo "== ncu full: promote kernel (conv4-sized launches)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_promote -s 40 -c 3 -f -o $out/r2_ncu_promote \
  $py bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --no-backbone --precision f16x3 > /dev/null 2>&1
echo "rc=$?"
ncu -i $out/r2_ncu_promote.ncu-rep --page raw --csv 2>/dev/null | $py -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=[i for i,h in enumerate(hdr) if any(k in h for k in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct','sm__warps_active.avg.pct','launch__registers_per_thread','launch__grid_size','sm__throughput.avg.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_bytes.sum '))]
for r in rows[:2]+rows[2:]:
    print(' | '.join(r[i][:60] for i in want))
" > $out/r2_ncu_promote.txt 2>&1
head -8 $out/r2_ncu_promote.txt
echo "== backbone micro-benchmark under ncu (tensor pipe per kernel)"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  --csv --log-file $out/r2_backbone_ncu.csv $py bench.py --config backbone > /dev/null 2>&1
echo "rc=$?"
