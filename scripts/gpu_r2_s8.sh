# Early accumulator release in the s2d epilogue (default build) and the experiment build with eight
# stem accumulators over two main ones (libopenglottal_b200_s8.so): tests, then a same-call A/B
mkdir -p gpurun_out
S8=$PWD/openglottal_b200/lib/libopenglottal_b200_s8.so
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_layers.py -m gpu -x -q 2>&1 | tail -3
OGL_LIB=$S8 timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "stem or schedules or north_star" 2>&1 | tail -3
: > gpurun_out/r2_exp_s8.jsonl
for rep in 1 2; do
  timeout 200 python scripts/layer_times.py 512 4 "slots4 rep=$rep" >> gpurun_out/r2_exp_s8.jsonl 2>> gpurun_out/r2_exp_s8.err
  OGL_LIB=$S8 timeout 200 python scripts/layer_times.py 512 4 "slots8 rep=$rep" >> gpurun_out/r2_exp_s8.jsonl 2>> gpurun_out/r2_exp_s8.err
done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_s8.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in L.items() if 'downs.0' in k or 'ups.7' in k})
PY
OGL_LIB=$S8 timeout 120 python scripts/stem_trace.py 128 2>&1 | tail -16
