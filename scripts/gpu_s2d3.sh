mkdir -p gpurun_out
rm -f gpurun_out/exp5.jsonl gpurun_out/exp5.err
timeout 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu.log | head -20; exit 0; fi
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp5.jsonl 2>> gpurun_out/exp5.err; }
run OGL_S2D=1
run OGL_S2D=1 OGL_DBG=4
run OGL_S2D=1 OGL_DBG=2
run OGL_S2D=1 OGL_DBG=1
run OGL_S2D=1 OGL_DBG=5
run OGL_S2D=1 OGL_S2D_SLOTS=3
run OGL_S2D=1 OGL_S2D_SLOTS=4
python scripts/show_exp.py gpurun_out/exp5.jsonl; tail -3 gpurun_out/exp5.err
