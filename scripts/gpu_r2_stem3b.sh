# s2d kernels: issuers wait for all stages of a tile first (OGL_S2D_WAITALL) -- tests, trace, A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_layers.py -m gpu -x -q > gpurun_out/r2_stem3_tests.log 2>&1; rc=$?; echo "model+layer tests rc=$rc"; tail -4 gpurun_out/r2_stem3_tests.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2_stem3_tests.log; exit 0; fi
timeout 120 python scripts/stem_trace.py 128 2>&1 | tail -22
: > gpurun_out/r2_exp_waitall.jsonl
for cfg in "OGL_FUSE_STEM=3" "OGL_FUSE_STEM=2" "OGL_FUSE_STEM=3 OGL_STEM_LO=0" "OGL_FUSE_STEM=3" "OGL_FUSE_STEM=2"; do
  env $cfg timeout 200 python scripts/layer_times.py 512 4 "$cfg" >> gpurun_out/r2_exp_waitall.jsonl 2>> gpurun_out/r2_exp_waitall.err
done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_waitall.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in L.items() if 'downs.0' in k or 'ups.7' in k})
PY
