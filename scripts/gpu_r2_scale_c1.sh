# The headline configuration (configs[1], weak scaling) at N = 1, 2, 4, 8 on one 8-GPU box with the
# final build, launched the way the driver launches bench.py; plus file -> features on 8 GPUs.
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/r2_scale_v3_c1_n1.json 2> gpurun_out/r2_scale_v3_c1_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --no-cpu-baseline > gpurun_out/r2_scale_v3_c1_n$n.json 2> gpurun_out/r2_scale_v3_c1_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_scale_v3_c1_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'sm', d['clocks']['sm_mhz'])
except Exception as e:
    print('N=$n no line', e)
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/dist_extract_check.py 200000 2> gpurun_out/r2_dist_extract_n8.err | tail -1 | tee gpurun_out/r2_dist_extract_n8.json
nproc
