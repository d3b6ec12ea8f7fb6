# upcat_tc (composed ConvTranspose2d at levels 1-3): layer tests first (short timeout: a hang in a
# new kernel must not take the box), then the model tests, then a same-call A/B of the two decoders.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -x -q -s -k upcat > gpurun_out/r2_upcat_layers.log 2>&1; echo "upcat layers rc=$?"; tail -15 gpurun_out/r2_upcat_layers.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -s > gpurun_out/r2_upcat_model.log 2>&1; echo "model rc=$?"; tail -5 gpurun_out/r2_upcat_model.log
for rep in 1 2; do
  for c in 1 0; do
    OGL_COMPOSE=$c timeout 300 python bench.py --no-cpu-baseline --layers-out gpurun_out/r2_layers_compose$c.json > gpurun_out/r2_bench_compose${c}_$rep.json 2> gpurun_out/r2_bench_compose${c}_$rep.err; echo "compose=$c rep $rep rc=$?"
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_compose*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'sm', d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'unreadable', e)
for c in (1, 0):
    try:
        print('compose', c)
        for l in json.load(open(f'gpurun_out/r2_layers_compose{c}.json')): print('  %-36s %8.4f ms %8.1f TF' % (l['layer'], l['ms'], l['tflops'] or 0))
    except Exception as e:
        print(e)
PY
