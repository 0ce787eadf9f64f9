cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - <<PY
import json
d=json.load(open('$1')); r=d['roofline']
print('N=%d value %.0f e2e %.0f ms/step %.1f dec_ms %.1f flow_ms %.1f text_ms %.1f frac %.4f clocks %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], r['dec_ms'], r['flow_ms'], r['text_ms'], r['frac'], d['clocks']))
PY
}
timeout 600 python bench.py --utts 2048 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n1s.json; show gpurun_out/bench_n1s.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --utts 2048 --steps 2 --warmup 2 > gpurun_out/bench_n2.log 2>&1
grep '^{' gpurun_out/bench_n2.log | tail -1 > gpurun_out/bench_n2.json; show gpurun_out/bench_n2.json
