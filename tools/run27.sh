cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 148 74 37; do
timeout 300 python tools/probe_conv.py 131072 $n > gpurun_out/probe_conv_sms$n.log 2>&1
echo "== num_sms $n"; grep -A4 "launch 2:\|launch 44:\|launch 49:" gpurun_out/probe_conv_sms$n.log | cut -c1-330
done
