cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv.log 2>&1
tail -3 gpurun_out/probe_conv.log
