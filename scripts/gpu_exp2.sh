mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp2.jsonl 2>> gpurun_out/exp2.err; }
rm -f gpurun_out/exp2.jsonl gpurun_out/exp2.err
run OGL_S128=1
run OGL_S128=2
run OGL_S128=1 OGL_DBG=4
run OGL_S128=1 OGL_DBG=16
wc -l gpurun_out/exp2.jsonl; tail -3 gpurun_out/exp2.err
