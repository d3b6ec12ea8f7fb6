mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -q -m gpu --timeout 120 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/rc.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'roofline frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
for l in json.load(open('gpurun_out/layers_n1.json')): print('%-24s %8.3f ms %8.1f TF' % (l['layer'], l['ms'], l['tflops'] or 0))
PY
tail -5 gpurun_out/bench.err
