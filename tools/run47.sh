cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rel_attention' -s 2 -c 1 -o gpurun_out/attn_mma -f python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out | grep attn
