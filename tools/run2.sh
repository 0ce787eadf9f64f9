set -x
cd $GRAFT_REPO_ROOT
nproc; free -g | head -2
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --utts 512 --steps 2 --warmup 1 --cpu-sample 8 2>&1 | tail -3 | tee gpurun_out/bench_512.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --utts 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout 600 python bench.py --impl reference --utts 512 --steps 1 --warmup 1 --cpu-sample 8 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
