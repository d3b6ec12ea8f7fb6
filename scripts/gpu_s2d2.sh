mkdir -p gpurun_out
rm -f gpurun_out/exp4.jsonl gpurun_out/exp4.err
timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_model.py -q -m gpu --timeout 100 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp4.jsonl 2>> gpurun_out/exp4.err; }
run OGL_S2D=1
run OGL_S2D=1 OGL_DBG=32
run OGL_S2D=1 OGL_DBG=36
run OGL_S2D=0 OGL_DBG=32
run OGL_S2D=0 OGL_DBG=36
run OGL_S2D=1 OGL_DBG=4
run OGL_S2D=1 OGL_S2D_SLOTS=3
python scripts/show_exp.py gpurun_out/exp4.jsonl; tail -3 gpurun_out/exp4.err
python scripts/profile_forward.py 128 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:s2d_tc_kernel -s 3 -c 3 -o gpurun_out/prof_s2d_v1 -f python scripts/profile_forward.py 128 > gpurun_out/ncu3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu3.log
