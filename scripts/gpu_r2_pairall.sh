mkdir -p gpurun_out
: > gpurun_out/r2_exp_s2d_rings.jsonl
for cfg in "OGL_DUAL_BELOW=0" "OGL_DUAL_BELOW=1 OGL_S2D_SLOTS_HEAD=6" "OGL_DUAL_BELOW=0" "OGL_DUAL_BELOW=1 OGL_S2D_SLOTS_HEAD=6" "OGL_S2D_SLOTS_HEAD=5" "OGL_S2D_SLOTS_HEAD=3"; do
  env $cfg timeout 200 python scripts/layer_times.py 512 4 "$cfg" >> gpurun_out/r2_exp_s2d_rings.jsonl 2>> gpurun_out/r2_exp_s2d_rings.err
done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_s2d_rings.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in L.items() if 'downs.0' in k or 'ups.7' in k})
PY
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_layers.py -m gpu -x -q 2>&1 | tail -3
