cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv_r01d.log 2>&1
timeout 300 python tools/probe_r01.py > gpurun_out/probe_mrf3_r01d.log 2>&1
tail -5 gpurun_out/probe_mrf3_r01d.log
