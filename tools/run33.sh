cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --utts 1024 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('batches', d['config']['device_batches'], 'chunk', d['config']['chunk_frames'], 'value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f dec TF %.1f frac %.4f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['achieved'], r['frac'], d['gpu_launches']))
" | tee -a gpurun_out/bench_v8.log
timeout 300 python tools/probe_conv.py > gpurun_out/probe_conv5.log 2>&1
grep -A2 "===" gpurun_out/probe_conv5.log | cut -c1-330
