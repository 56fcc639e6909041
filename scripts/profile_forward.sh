#!/bin/bash
# ncu evidence for the headline workload (run under gpurun, one GPU).  Usage: scripts/profile_forward.sh <tag>
# 1. plain run (must exit 0 without ncu)   2. launch list of the same command   3. one `--set full` capture of the fused
# block kernel.  Outputs land in gpurun_out/; summaries are copied into profiles/ by hand (see profiles/README.md).
set -e
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-train"
$CMD > gpurun_out/plain_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:resblock|dense2|chain|ncl|nlc|leaky|softmax|featur|taps' -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD \
  > gpurun_out/ncu_l_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resblock2_kernel -s 25 -c 2 -f \
  -o gpurun_out/prof_${TAG}_resblock2 $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}_resblock2.ncu-rep --page raw --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__cluster_size,launch__shared_mem_per_block_dynamic,sm__cycles_elapsed.avg.per_second,sm__warps_active.avg.pct_of_peak_sustained_active \
  > gpurun_out/ncu_full_${TAG}_resblock2.csv 2>/dev/null || true
tail -3 gpurun_out/ncu_full_${TAG}_resblock2.csv
