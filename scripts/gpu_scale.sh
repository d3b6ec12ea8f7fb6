# usage (under gpurun --gpus 8): gpu_scale.sh "1 2 4 8"  -- bench.py at each N back to back on one
# box, launched the way the driver launches it; lines go to gpurun_out/bench_n<N>.json
mkdir -p gpurun_out
for n in ${1:-1 2 4 8}; do
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks'])
except Exception as e:
    print('N=$n no line', e)
PY
done
