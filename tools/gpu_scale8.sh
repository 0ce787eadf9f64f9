# 8-GPU weak-scaling check (the driver's own SCALE run is the one that counts): bench under torchrun on one box
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --utts 2048 --steps 2 --warmup 2 > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json')); r=d['roofline']
print('N=%d value %.0f e2e %.0f ms/step %.1f busy %.1f frac %.4f clocks %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['device_busy_ms_per_step'], r['frac'], d['clocks']))
PY
tail -3 gpurun_out/bench_n$N.log | cut -c1-300
