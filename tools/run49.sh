cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "ln_rpw=1" "ln_rpw=2" "ln_rpw=4" "ln_rpw=1"; do
timeout 600 python bench.py --utts 2048 --steps 2 --warmup 2 --no-cpu-baseline --opt $cfg 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_ab.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json')); r=d['roofline']
print('$cfg', 'value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f dec TF %.1f frac %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['achieved'], r['frac']))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rel_attention' -s 2 -c 1 -o gpurun_out/attn_mma2 -f python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn2.log 2>&1
