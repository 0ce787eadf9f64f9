# is the first CUDA process on a fresh box slower between launches?  bench twice, nothing before it
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --utts 2048 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_fp$i.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_fp$i.json')); r=d['roofline']
print('run $i value %.0f e2e %.0f ms/step %.1f busy %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f frac %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['device_busy_ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['frac']))
PY
done
