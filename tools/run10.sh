cd $GRAFT_REPO_ROOT
B="python bench.py --utts 160 --steps 1 --warmup 1 --no-cpu-baseline --chunk-frames 32768"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mrf_tc -s 2 -c 2 -o gpurun_out/prof_mrf_r01 $B > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 45 -c 4 -o gpurun_out/prof_convtc_flow_r01 $B > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 77 -c 13 -o gpurun_out/prof_convtc_dec_r01 $B > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
