# same-box A/B of the fused-stem modes with bench.py itself (default 100 timed steps)
mkdir -p gpurun_out
for m in 1 2 1 2 1 2; do
OGL_FUSE_STEM=$m timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_stem_ab.json 2> gpurun_out/bench_stem_ab.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_stem_ab.json').read().strip().splitlines()[-1])
print('mode $m value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'MHz', d['clocks']['sm_mhz'], 'W', d['clocks'].get('mean_w'))
PY
done
