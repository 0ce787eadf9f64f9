# A/B of an engine option on one box (two runs per arm, alternating): gpu_ab_opt.sh <tag> <key> <valueA> <valueB> [bench flags]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
KEY=$1; shift
VA=$1; shift
VB=$1; shift
for V in $VA $VB $VA $VB; do
timeout 600 python bench.py --utts 2048 --steps 2 --warmup 3 --no-cpu-baseline --no-strong --opt $KEY=$V "$@" 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_ab.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json')); r=d['roofline']
print('[$KEY=$V]', 'value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f frac %.4f k32 %.3f k64 %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['frac'], r['kernels'][0]['frac'], r['kernels'][1]['frac']))
PY
done 2>&1 | tee gpurun_out/ab_$TAG.log
