mkdir -p gpurun_out
: > gpurun_out/r2_exp_misc.jsonl
for cfg in "OGL_CG_MINKB=2" "OGL_CG_MINKB=1" "OGL_NA=4" "OGL_NW=4" "OGL_UP_NA=3 OGL_UP_NW=6" "OGL_CG_MINKB=2"; do
  env $cfg timeout 200 python scripts/layer_times.py 512 4 "$cfg" >> gpurun_out/r2_exp_misc.jsonl 2>> gpurun_out/r2_exp_misc.err
done
python - <<'PY'
import json
rows = [json.loads(l) for l in open('gpurun_out/r2_exp_misc.jsonl')]
names = list(rows[0]['layers'])
print('%-36s' % 'launch', *['%14s' % r['tag'][-14:] for r in rows])
for n in names:
    print('%-36s' % n, *['%14.4f' % r['layers'].get(n, float('nan')) for r in rows])
print('%-36s' % 'step', *['%14.3f' % r['ms_step'] for r in rows])
print('%-36s' % 'sm MHz', *['%14d' % r['clocks']['sm_mhz'] for r in rows])
PY
