set -e
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:"xent_fwd_reg|softmax_fwd_reg|layernorm_fwd_reg|argmax_reg_wide" -s 8 -c 4 -o gpurun_out/r2s2_bw python scripts/bench_bandwidth.py > gpurun_out/r2s2_ncu_bw.log 2>&1 || tail -5 gpurun_out/r2s2_ncu_bw.log
