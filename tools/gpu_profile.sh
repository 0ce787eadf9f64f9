# ncu captures behind profiles/: launch list of one short bench run + `--set full` of the dominant kernels.
# usage (from this container):  gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh <tag>'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-x}
B="python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_list_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | tee gpurun_out/launch_shares_$TAG.txt
# the two fused MRF kernels (first full 131072-frame chunk of the timed step), then every k_conv_tc of the same chunk's decoder
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mrf3' -s 8 -c 2 -o gpurun_out/mrf_full_$TAG -f $B > gpurun_out/ncu_mrf_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' -s ${CONV_SKIP:-75} -c 9 -o gpurun_out/conv_dec_$TAG -f $B > gpurun_out/ncu_conv_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rel_attention' -s 2 -c 1 -o gpurun_out/attn_$TAG -f $B > gpurun_out/ncu_attn_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
