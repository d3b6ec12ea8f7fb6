mkdir -p gpurun_out
rm -f gpurun_out/s2d_*.log gpurun_out/exp3.jsonl gpurun_out/exp3.err
timeout 300 python -m pytest tests/test_gpu_layers.py -k s2d -q -m gpu --timeout 100 --timeout-method=thread -x -s > gpurun_out/s2d_layers.log 2>&1; rc=$?; echo "s2d layers rc=$rc"; tail -15 gpurun_out/s2d_layers.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp3.jsonl 2>> gpurun_out/exp3.err; }
run OGL_S2D=1
run OGL_S2D=0
run OGL_S2D=1 OGL_DBG=1
run OGL_S2D=1 OGL_DBG=4
run OGL_S2D=1 OGL_DBG=2
python scripts/show_exp.py gpurun_out/exp3.jsonl; tail -3 gpurun_out/exp3.err
