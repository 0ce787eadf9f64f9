# ncu evidence for the memory-bound kernels (north_star: achieved HBM GB/s against B200 peak): gpu_membound.sh <tag>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-x}
export PASSES=2
K='regex:k_layernorm|k_expand_sample|k_spline_inverse|k_durations|k_frame_index|k_absmax|k_to_int16|k_noise_dp|k_cf_pre|k_ea_logw|k_embed|k_row_pos'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --csv \
    --log-file gpurun_out/membound_$TAG.csv python tools/membound_workload.py > gpurun_out/membound_$TAG.out 2>&1
grep '^{' gpurun_out/membound_$TAG.out | tail -1 > gpurun_out/membound_sizes_$TAG.json
python tools/ncu_membound.py gpurun_out/membound_$TAG.csv gpurun_out/membound_sizes_$TAG.json | tee gpurun_out/membound_$TAG.txt
