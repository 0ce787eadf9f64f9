cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv4.log 2>&1
cut -c1-420 gpurun_out/probe_conv4.log | grep -A3 "==="
