# Round-2: the GPU test suite, then one bench line per BASELINE config (N = 1) + the cuDNN context arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log
for c in 0 1 2 3; do
  timeout 600 python bench.py --config $c --layers-out gpurun_out/r2_layers_c$c.json > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err; echo "config $c rc=$?"
done
timeout 300 python bench.py --config 0 --no-graph --no-cpu-baseline > gpurun_out/r2_bench_c0_nograph.json 2> gpurun_out/r2_bench_c0_nograph.err; echo "config 0 nograph rc=$?"
timeout 900 python bench.py --config 4 --frames 200000 > gpurun_out/r2_bench_c4_200k.json 2> gpurun_out/r2_bench_c4_200k.err; echo "config 4 (200k) rc=$?"
for c in 0 1 2; do
  timeout 600 python bench.py --impl cudnn --config $c > gpurun_out/r2_cudnn_c$c.json 2> gpurun_out/r2_cudnn_c$c.err; echo "cudnn $c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_ref_c1.json 2> gpurun_out/r2_ref_c1.err; echo "reference rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_c*.json') + glob.glob('gpurun_out/r2_cudnn_c*.json') + glob.glob('gpurun_out/r2_ref_c*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d.get('value', 0)), 'e2e', round(d.get('e2e', {}).get('value', 0)), 'ms/step', round(d.get('ms_per_step', 0), 3),
              'frac', round(d.get('roofline', {}).get('frac', 0), 3), 'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
