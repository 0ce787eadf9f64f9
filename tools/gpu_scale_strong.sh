# strong scaling only, both gather variants: gpu_scale_strong.sh <tag> <N>
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; N=$2
for V in "direct" "memcpy --gather-memcpy" "direct64k --chunk-frames 65536"; do
  set -- $V; name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --scaling strong --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/strong_${name}_n${N}_$TAG.err | grep '^{' | tail -1 > gpurun_out/bench_strong_${name}_n${N}_$TAG.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_strong_${name}_n${N}_$TAG.json'))
    print('[strong $name N=$N]', 'value %.0f e2e %.0f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['e2e'].get('passes_s'))
except Exception as e:
    print('[strong $name N=$N] failed', e); print(open('gpurun_out/strong_${name}_n${N}_$TAG.err').read()[-2000:])
PY
done
