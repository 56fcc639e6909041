#!/bin/bash
# End-of-session evidence on one GPU (run under gpurun): full GPU test suite, smoke(), the plain default bench line, then
# (only after the plain run exited 0) the launch list of one forward and one `--set full` capture of the block kernel.
TAG=${1:-r2s2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err || { tail -5 gpurun_out/${TAG}_bench_default.err; exit 1; }
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-train --no-stock --no-longread"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:resblock|dense2|chain|ncl|nlc|leaky|softmax|featur|taps|embed' -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resblock3_kernel -s 25 -c 1 -f -o gpurun_out/${TAG}_resblock3 $CMD \
  > gpurun_out/${TAG}_ncu_f.log 2>&1
ncu -i gpurun_out/${TAG}_resblock3.ncu-rep --page raw --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__cluster_size,launch__shared_mem_per_block_dynamic,sm__cycles_elapsed.avg.per_second,smsp__inst_executed.sum \
  > gpurun_out/${TAG}_ncu_full_resblock3.csv 2>/dev/null || true
tail -2 gpurun_out/${TAG}_ncu_full_resblock3.csv
ls gpurun_out | head -30
