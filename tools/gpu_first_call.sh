#!/bin/bash
# First gpurun call of a round, everything in ONE box session (run from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_first_call.sh r2'
# 1. the GPU tests that were written after round 1's GPU budget was spent (tests/test_zz_next_rows_gpu.py),
# 2. the whole GPU suite, 3. the default bench line, 4. the serialised launch list of one eager step,
# 5. stand-alone timings (tools/bench_proposal_kernels.py) and `ncu --set full` of the non-GEMM kernels DESIGN.md
#    section 7 ranks first (ROIAlign, radix sort, NMS bitmask + scan).
# Every stage has its own timeout and writes under gpurun_out/<tag>_*; a failing stage does not stop the others.
tag=${1:-rN}
out=gpurun_out
mkdir -p $out
py=python
echo "== new tests"; timeout 900 $py -m pytest tests/test_zz_next_rows_gpu.py -q -rxX --runxfail > $out/${tag}_new_tests.log 2>&1; echo "rc=$?"; tail -5 $out/${tag}_new_tests.log
echo "== gpu suite"; timeout 1500 $py -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "rc=$?"; tail -3 $out/${tag}_gpu_tests.log
echo "== bench"; timeout 600 $py bench.py --steps 20 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err; echo "rc=$?"; cut -c1-400 $out/${tag}_bench_1gpu.json
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
  $py bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > /dev/null 2>&1
echo "rc=$?"; $py tools/agg_launches.py $out/${tag}_launches.csv 7 > $out/${tag}_launches_step.txt 2>&1; head -12 $out/${tag}_launches_step.txt
echo "== stand-alone kernel timings"; timeout 300 $py tools/bench_proposal_kernels.py > $out/${tag}_proposal_kernels.txt 2>&1; echo "rc=$?"; cat $out/${tag}_proposal_kernels.txt
# ncu --set full on the stand-alone tool (no model build: seconds per capture); -s skips the warm-up launches
for spec in "roi_align_roi_kernel roialign" "segmented_radix_sort_kernel proposals" "nms_scan_kernel proposals" "nms_bitmask_kernel proposals"; do
  set -- $spec; k=$1; what=$2
  echo "== ncu full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/${tag}_ncu_$k \
    $py tools/bench_proposal_kernels.py $what > /dev/null 2>&1
  echo "rc=$?"
  ncu -i $out/${tag}_ncu_$k.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|DRAM Throughput|L2 Cache Throughput|L1/TEX Cache Throughput|Achieved Occupancy|Registers Per|Theoretical Occupancy|Executed Ipc Active|No Eligible|Issued Warp|L2 Hit|Mem Busy|Max Bandwidth" | cut -c1-200 > $out/${tag}_ncu_$k.txt
  head -16 $out/${tag}_ncu_$k.txt
done
