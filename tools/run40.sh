cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_full.log 2>&1
tail -1 gpurun_out/bench_full.log | grep '^{' > gpurun_out/bench_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']
print('value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f dec TF %.1f frac %.4f launches %d cpu %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['achieved'], r['frac'], d['gpu_launches'], d.get('cpu_baseline',{}).get('value')))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launch_shares.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mrf3|k_mrf2' -s 8 -c 4 -o gpurun_out/mrf_full -f python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mrf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc' -s 200 -c 12 -o gpurun_out/conv_full -f python bench.py --utts 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out
