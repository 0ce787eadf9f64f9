# strong + weak scaling on N GPUs of one box: gpu_scale.sh <tag> <N>   (run under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; N=$2
run() { # <name> <flags...>
  name=$1; shift
  if [ "$N" = "1" ]; then timeout 900 python bench.py --gpus 1 "$@" 2> gpurun_out/scale_${name}_n${N}_$TAG.err | grep '^{' | tail -1 > gpurun_out/bench_${name}_n${N}_$TAG.json
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2> gpurun_out/scale_${name}_n${N}_$TAG.err | grep '^{' | tail -1 > gpurun_out/bench_${name}_n${N}_$TAG.json; fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${name}_n${N}_$TAG.json'))
    print('[$name N=$N]', 'value %.0f e2e %.0f ms/step %.1f frac %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']), d['e2e'].get('passes_s'), d.get('strong_scaling'))
except Exception as e:
    print('[$name N=$N] failed', e); print(open('gpurun_out/scale_${name}_n${N}_$TAG.err').read()[-2500:])
PY
}
run strong --scaling strong --steps 3 --warmup 3 --no-cpu-baseline
run weak --steps 3 --warmup 3 --no-cpu-baseline
