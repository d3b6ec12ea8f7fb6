mkdir -p gpurun_out
rm -f gpurun_out/exp_sweep.jsonl gpurun_out/exp_sweep.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_sweep.jsonl 2>> gpurun_out/exp_sweep.err; }
run OGL_X=base
run OGL_CG_MINKB=1
run OGL_S2D_PAIR_ALL=1
run OGL_S128=1
run OGL_SPLIT=0
run OGL_ST=1
run OGL_NA=2
run OGL_X=base
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/exp_sweep.jsonl')]
names=list(rows[0]['layers'])
print('%-22s %6s %5s '%('env','ms','MHz')+' '.join('%5s'%n.replace('downs.','d').replace('.net.','c').replace('bottleneck','b').replace('ups.','u').replace('(convT)','T').replace('(cat)','').replace('+pool','p').replace('stem+','s')[:5] for n in names))
for r in rows:
    e=' '.join(f'{k[4:]}={v}' for k,v in r['env'].items())
    print('%-22s %6.2f %5s '%(e[:22],r['ms_step'],r['clocks']['sm_mhz'])+' '.join('%5.2f'%r['layers'].get(n,float('nan')) for n in names))
PY
tail -2 gpurun_out/exp_sweep.err
