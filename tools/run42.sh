cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' -s 78 -c 8 -o gpurun_out/conv_stage1 -f python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_conv1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' -s 39 -c 3 -o gpurun_out/conv_flow -f python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_conv2.log 2>&1
ls -la gpurun_out
