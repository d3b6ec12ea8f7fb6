mkdir -p gpurun_out
rm -f gpurun_out/exp_pp.jsonl gpurun_out/exp_pp.err
timeout 600 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
OGL_PINGPONG=0 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_layers.py -q -m gpu --timeout 300 --timeout-method=thread -x > gpurun_out/pytest_gpu_pp0.log 2>&1; echo "pytest(pp0) rc=$?"; tail -2 gpurun_out/pytest_gpu_pp0.log
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_pp.jsonl 2>> gpurun_out/exp_pp.err; }
run OGL_PINGPONG=0
run OGL_PINGPONG=1
run OGL_PINGPONG=0
run OGL_PINGPONG=1
run OGL_PINGPONG=0
run OGL_PINGPONG=1
python scripts/show_exp.py gpurun_out/exp_pp.jsonl; tail -3 gpurun_out/exp_pp.err
