mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -k "fused_stem" -q -m gpu --timeout 100 --timeout-method=thread -x -s > gpurun_out/stem.log 2>&1; rc=$?; echo "fused stem rc=$rc"; tail -12 gpurun_out/stem.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_gpu.log
rm -f gpurun_out/exp_stem.jsonl gpurun_out/exp_stem.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp_stem.jsonl 2>> gpurun_out/exp_stem.err; }
run OGL_FUSE_STEM=0
run OGL_FUSE_STEM=1
run OGL_FUSE_STEM=0
run OGL_FUSE_STEM=1
python scripts/show_exp.py gpurun_out/exp_stem.jsonl; tail -3 gpurun_out/exp_stem.err
