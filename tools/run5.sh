cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mrf_tc -s 4 -c 2 -o gpurun_out/prof_mrf python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 60 -c 6 -o gpurun_out/prof_convtc python bench.py --utts 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench4.log 2>&1
ls -la gpurun_out/
