cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv3.log 2>&1
grep -A3 "launch 2:\|launch 3:\|launch 44:\|launch 50:" gpurun_out/probe_conv3.log | cut -c1-400
