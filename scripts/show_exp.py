"""Tabulates the JSON lines scripts/layer_times.py wrote (one row per run, one column per launch)."""
import json, sys
rows=[json.loads(l) for l in open(sys.argv[1])]
names=[]
for r in rows:
    for n in r['layers']:
        if n not in names: names.append(n)
def short(n):
    return n.replace('downs.','d').replace('.net.','c').replace('bottleneck','b').replace('ups.','u').replace('(convT)','T').replace('(cat)','')[:7]
print('%-40s %7s %5s '%('env','ms','MHz')+' '.join('%7s'%short(n) for n in names))
for r in rows:
    e=' '.join(f'{k[4:]}={v.split("/")[-1]}' for k,v in r['env'].items())
    print('%-40s %7.2f %5s '%(e[:40],r['ms_step'],str((r.get('clocks') or {}).get('sm_mhz')))+' '.join(('%7.2f'%r['layers'][n]) if n in r['layers'] else '      -' for n in names))
