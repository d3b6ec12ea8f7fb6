# Final check of the 16-warp stem (fuse_stem 3): whole GPU suite, bench.py A/B against mode 2, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v2.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_pytest_v2.log
for m in 3 2 3 2; do
  OGL_FUSE_STEM=$m timeout 300 python bench.py --no-cpu-baseline --layers-out gpurun_out/r2_layers_stem$m.json > gpurun_out/r2_bench_stem$m.json 2> gpurun_out/r2_bench_stem$m.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_stem$m.json').read().strip().splitlines()[-1]); print('stem=$m value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'sm', d['clocks']['sm_mhz'])"
done
timeout 120 python scripts/stem_trace.py 128 > gpurun_out/r2_stem_trace.txt 2>&1; head -8 gpurun_out/r2_stem_trace.txt
timeout 600 python scripts/ingest_bench.py 40000 MJPG > gpurun_out/r2_ingest_mjpg.json 2> gpurun_out/r2_ingest_mjpg.err; echo "ingest rc=$?"; cat gpurun_out/r2_ingest_mjpg.json; tail -3 gpurun_out/r2_ingest_mjpg.err
