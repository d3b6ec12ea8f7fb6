mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -q -m gpu --timeout 100 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
rm -f gpurun_out/exp_diag.jsonl gpurun_out/exp_diag.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 3 "$*" >> gpurun_out/exp_diag.jsonl 2>> gpurun_out/exp_diag.err; }
run OGL_FUSE_STEM=0 OGL_CG=1
run OGL_FUSE_STEM=0 OGL_CG=1 OGL_EXPERIMENT=1 OGL_DBG=12
run OGL_FUSE_STEM=0 OGL_CG=1 OGL_EXPERIMENT=1 OGL_DBG=13
run OGL_FUSE_STEM=1 OGL_CG=2
run OGL_FUSE_STEM=1 OGL_CG=2
python - <<'PY'
import json
for l in open('gpurun_out/exp_diag.jsonl'):
    d=json.loads(l); L=d['layers']
    g=lambda k: L.get(k, float('nan'))
    print('%-40s %6.2f MHz %s  d0c3 %.2f/%.2f  u7c0 %.2f  u7c3 %.2f  u5c0 %.2f' % (' '.join(f"{k[4:]}={v}" for k,v in d['env'].items()), d['ms_step'], d['clocks']['sm_mhz'], g('downs.0.net.3+pool'), g('stem+downs.0.net.3+pool'), g('ups.6(convT)+ups.7.net.0(cat)'), g('ups.7.net.3+head'), g('ups.5.net.0(cat)')))
PY
tail -3 gpurun_out/exp_diag.err
