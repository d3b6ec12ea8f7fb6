# Round-2 scaling evidence on one 8-GPU box (gpurun --gpus 8), launched the way the driver launches
# bench.py: config 4 (10^6-frame clip, STRONG scaling, feature step) at N = 1, 2, 4, 8; config 2
# (512x256 native, 25 000 frames per rank) at N = 8 and 1; config 1 (the headline) at N = 8.
mkdir -p gpurun_out
run() {  # run <tag> <n> <args...>
  tag=$1; n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
  fi
  echo "$tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_scale_$tag.json').read().strip().splitlines()[-1])
    print('$tag value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'steps', d['steps'], 'frac', round(d['roofline']['frac'],3), 'scaling', d['scaling'], 'feature_ms', d.get('feature_step_ms'), 'sm', d['clocks']['sm_mhz'])
except Exception as e:
    print('$tag no line', e)
PY
}
run c4_n1 1 --config 4 --frames 1000000
for n in 2 4 8; do run c4_n$n $n --config 4 --frames 1000000 --no-cpu-baseline; done
run c2_n8 8 --config 2 --no-cpu-baseline
run c1_n8 8 --config 1 --no-cpu-baseline
run c3_n8 8 --config 3 --no-cpu-baseline
