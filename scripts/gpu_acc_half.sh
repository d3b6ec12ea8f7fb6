# per-half accumulator-empty barriers (OGL_ACC_HALF): GPU suite, then A/B on one box
mkdir -p gpurun_out
rm -f gpurun_out/exp_half.jsonl gpurun_out/exp_half.err
timeout 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED|Timeout" gpurun_out/pytest_gpu.log | head; exit 0; fi
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_half.jsonl 2>> gpurun_out/exp_half.err; }
run OGL_ACC_HALF=0
run OGL_ACC_HALF=1
run OGL_ACC_HALF=0
run OGL_ACC_HALF=1
python scripts/show_exp.py gpurun_out/exp_half.jsonl | cut -c1-230; tail -3 gpurun_out/exp_half.err
