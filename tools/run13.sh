cd $GRAFT_REPO_ROOT
B="python bench.py --utts 160 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mrf2_tc -s 2 -c 2 -o gpurun_out/prof_mrf2_r01 $B > gpurun_out/ncu_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
