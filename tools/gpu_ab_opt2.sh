# like gpu_ab_opt.sh, with free-form bench flags per arm: gpu_ab_opt2.sh <tag> "<flags A>" "<flags B>" [common flags]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
FA=$1; shift
FB=$1; shift
for F in "$FA" "$FB" "$FA" "$FB"; do
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong $F "$@" 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_ab.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json')); r=d['roofline']
print('[$F]', 'value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f frac %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['frac']))
PY
done 2>&1 | tee gpurun_out/ab_$TAG.log
