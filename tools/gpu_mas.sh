#!/bin/bash
# Monotonic alignment search on the GPU box: parity tests, the bench line, and an ncu pass for the kernels' DRAM traffic.
# usage: bash tools/gpu_mas.sh <tag>
tag=${1:-r02v}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mas.py -x -q 2>&1 | tail -4
timeout 120 python tools/bench_mas.py > gpurun_out/bench_mas_${tag}.json 2> gpurun_out/bench_mas_${tag}.err
cat gpurun_out/bench_mas_${tag}.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_mas -c 12 --csv \
    --log-file gpurun_out/ncu_mas_${tag}.csv python tools/bench_mas.py --iters 2 > /dev/null 2> gpurun_out/ncu_mas_${tag}.err
tail -8 gpurun_out/ncu_mas_${tag}.csv | cut -c1-220
