# upcat ring-depth sweep (same box, same call), then the full GPU suite and the per-config bench lines
mkdir -p gpurun_out
: > gpurun_out/r2_exp_upring.jsonl
for cfg in "0 0" "2 0" "2 6" "3 4"; do
  set -- $cfg
  OGL_UP_NA=$1 OGL_UP_NW=$2 timeout 200 python scripts/layer_times.py 512 5 "na=$1 nw=$2" >> gpurun_out/r2_exp_upring.jsonl 2>> gpurun_out/r2_exp_upring.err
done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_upring.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: v for k, v in L.items() if 'convT' in k})
PY
