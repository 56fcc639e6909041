set -e
cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:taps_fwd_kernel -s 6 -c 3 -o gpurun_out/r2s2_taps python scripts/bench_extra.py bytenet B=8 T=4096 > gpurun_out/r2s2_ncu_taps.log 2>&1 || tail -5 gpurun_out/r2s2_ncu_taps.log
ls -la gpurun_out/
