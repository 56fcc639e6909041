# ncu --set full of the second session's bandwidth kernels: 8 launches in all (2 each), bounded.  Run under gpurun.
cd $GRAFT_REPO_ROOT
python scripts/bw_once.py || exit 1
timeout 300 ncu --set full --clock-control none -k regex:"ncl_to_nlc_v3_kernel|avgpool_ncl_to_nlc_v3_kernel|softmax_fwd_reg|layernorm_fwd_reg" -c 8 -f -o gpurun_out/r2s2_bw2 python scripts/bw_once.py > gpurun_out/r2s2_ncu_bw2.log 2>&1
ncu -i gpurun_out/r2s2_bw2.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__cycles_elapsed.avg.per_second > gpurun_out/r2_ncu_full_bw_session2.csv 2>/dev/null
rm -f gpurun_out/r2s2_bw2.ncu-rep
cat gpurun_out/r2_ncu_full_bw_session2.csv | cut -c1-400
