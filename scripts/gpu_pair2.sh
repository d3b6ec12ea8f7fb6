mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py -k "s2d" -q -m gpu --timeout 60 --timeout-method=thread -x -s > gpurun_out/pair_layers.log 2>&1; rc=$?; echo "s2d layers rc=$rc"; tail -12 gpurun_out/pair_layers.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 0; fi
rm -f gpurun_out/exp_pair2.jsonl gpurun_out/exp_pair2.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp_pair2.jsonl 2>> gpurun_out/exp_pair2.err; }
run OGL_CG=1
run OGL_CG=2
run OGL_CG=1
run OGL_CG=2
python scripts/show_exp.py gpurun_out/exp_pair2.jsonl; tail -3 gpurun_out/exp_pair2.err
