# A new build against the previous one (openglottal_b200/lib/libopenglottal_b200_prev.so): whole GPU suite, then a
# same-call A/B against the previous build (openglottal_b200/lib/libopenglottal_b200_prev.so)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v3.log 2>&1; rc=$?; echo "all gpu tests rc=$rc"; tail -3 gpurun_out/r2_pytest_v3.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/r2_pytest_v3.log | head -20; fi
PREV=$PWD/openglottal_b200/lib/libopenglottal_b200_prev.so
: > gpurun_out/r2_exp_ring.jsonl
for rep in 1 2; do
  timeout 200 python scripts/layer_times.py 512 4 "new rep=$rep" >> gpurun_out/r2_exp_ring.jsonl 2>> gpurun_out/r2_exp_ring.err
  OGL_LIB=$PREV timeout 200 python scripts/layer_times.py 512 4 "prev rep=$rep" >> gpurun_out/r2_exp_ring.jsonl 2>> gpurun_out/r2_exp_ring.err
done
python - <<'PY'
import json
rows = [json.loads(l) for l in open('gpurun_out/r2_exp_ring.jsonl')]
names = list(rows[0]['layers'])
print('%-36s' % 'launch', *['%12s' % r['tag'] for r in rows])
for n in names:
    print('%-36s' % n, *['%12.4f' % r['layers'].get(n, float('nan')) for r in rows])
print('%-36s' % 'step', *['%12.3f' % r['ms_step'] for r in rows])
print('%-36s' % 'sm MHz', *['%12d' % r['clocks']['sm_mhz'] for r in rows])
PY
for v in new prev new prev new prev; do
  if [ $v = prev ]; then export OGL_LIB=$PREV; else unset OGL_LIB; fi
  timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_ring_$v.json 2> gpurun_out/r2_bench_ring_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_ring_$v.json').read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'sm', d['clocks']['sm_mhz'])"
done
