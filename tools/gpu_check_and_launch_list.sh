cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --utts 2048 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_$TAG.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json')); r=d['roofline']
print('value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f dec TF %.1f frac %.4f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['achieved'], r['frac'], d['gpu_launches']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | tee gpurun_out/launch_shares_$TAG.txt
