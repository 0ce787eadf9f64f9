# bench lines of every BASELINE.json config on one box: gpu_configs.sh <tag> [extra bench flags]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
for C in C1 C2 C3 C4; do
timeout 900 python bench.py --config $C --steps 5 --warmup 3 "$@" 2> gpurun_out/bench_${C}_$TAG.err | grep '^{' | tail -1 > gpurun_out/bench_${C}_$TAG.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${C}_$TAG.json')); r=d['roofline']
    print('[$C]', 'value %.0f e2e %.0f ms/step %.2f dec_ms %.1f flow_ms %.1f text_ms %.1f dec frac %.4f launches %d cpu %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['frac'], d['gpu_launches'], d.get('cpu_baseline', {}).get('value')), d.get('latency_ms', ''), d.get('spot_check', ''))
except Exception as e:
    print('[$C] failed', e); print(open('gpurun_out/bench_${C}_$TAG.err').read()[-1500:])
PY
done 2>&1 | tee gpurun_out/configs_$TAG.log
