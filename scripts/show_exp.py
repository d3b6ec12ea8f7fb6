import json, sys
rows=[json.loads(l) for l in open(sys.argv[1])]
names=list(rows[0]['layers'].keys())
short=[n.replace('downs.','d').replace('.net.','c').replace('bottleneck','b').replace('ups.','u')[:12] for n in names]
print('%-26s %7s '%('env','ms')+' '.join('%6s'%s[:6] for s in short))
for r in rows:
    e=' '.join(f'{k[4:]}={v}' for k,v in r['env'].items())
    print('%-26s %7.2f '%(e[:26],r['ms_step'])+' '.join('%6.2f'%r['layers'][n] for n in names))
