cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r01.json
cat gpurun_out/bench_r01.json | cut -c1-1500
tail -3 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --utts 160 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-300
B="python bench.py --utts 160 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mrf2_tc -s 2 -c 2 -o gpurun_out/prof_mrf2_r01 $B > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 60 -c 12 -o gpurun_out/prof_convtc_r01 $B > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/
