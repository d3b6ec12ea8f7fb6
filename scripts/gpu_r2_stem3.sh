# 16-warp tensor-core stem (fuse_stem = 3): stem tests under a short timeout (a hang in a new kernel
# must not take the box), then a same-call A/B of the per-launch times and of bench.py, modes 2 / 3.
mkdir -p gpurun_out
for b in 0 1; do OGL_STEM3_BFMT=$b CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/stem3_diag.py 2>&1 | tail -2; echo "diag bfmt=$b rc=$?"; done
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -s -k "stem or schedules or idempotent" > gpurun_out/r2_stem3_tests.log 2>&1; rc=$?; echo "stem tests rc=$rc"; tail -25 gpurun_out/r2_stem3_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
: > gpurun_out/r2_exp_stem3.jsonl
for rep in 1 2; do for m in 2 3; do
  OGL_FUSE_STEM=$m timeout 200 python scripts/layer_times.py 512 5 "stem=$m rep=$rep" >> gpurun_out/r2_exp_stem3.jsonl 2>> gpurun_out/r2_exp_stem3.err
done; done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_stem3.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in L.items() if 'stem' in k})
PY
for m in 3 2 3 2; do
  OGL_FUSE_STEM=$m timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_stem$m.json 2> gpurun_out/r2_bench_stem$m.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_stem$m.json').read().strip().splitlines()[-1]); print('stem=$m value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'sm', d['clocks']['sm_mhz'])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_stem3_pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_stem3_pytest_all.log
