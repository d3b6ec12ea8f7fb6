mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_sust.json 2> gpurun_out/bench_sust.err; echo "bench(default 100/20) rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench(20/3) rc=$?"
timeout 600 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench_400.json 2> gpurun_out/bench_400.err; echo "bench(400/20) rc=$?"
python - <<'PY'
import json
for f in ('sust','short','400'):
    d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
    print(f, d['steps'], d['warmup'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks'])
PY
