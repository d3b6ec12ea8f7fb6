# Round-2 first call: compose A/B (same box), then the whole GPU suite and one bench line per config.
mkdir -p gpurun_out
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
bash scripts/gpu_r2_configs.sh
