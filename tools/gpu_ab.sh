# A/B of bench.py configurations on one box: each argument is one quoted flag string
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "$@"; do
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline $cfg 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_ab.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json')); r=d['roofline']
print('[$cfg]', 'value %.0f e2e %.0f ms/step %.1f busy %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f frac %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['device_busy_ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['frac']))
PY
done
