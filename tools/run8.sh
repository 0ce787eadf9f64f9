cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mrf_tc -s 3 -c 1 -o gpurun_out/prof_mrf2 python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline --chunk-frames 32768 > gpurun_out/ncu_bench3.log 2>&1
ls -la gpurun_out/*.ncu-rep
