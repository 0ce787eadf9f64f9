cd $GRAFT_REPO_ROOT
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) 2>&1 | tail -6 | tee gpurun_out/bench_default.json
timeout 600 python bench.py --utts 512 --steps 2 --warmup 3 --no-cpu-baseline --chunk-frames 32768 2>&1 | tail -1 | tee gpurun_out/bench_512_cf32k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --utts 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -1 gpurun_out/ncu_bench2.log
