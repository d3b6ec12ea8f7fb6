# tensor-core stem: its own test under a short timeout first (a pipeline bug would hang), then the
# whole GPU suite with the tensor-core stem as the process default, then A/B of the stem modes
mkdir -p gpurun_out
rm -f gpurun_out/exp_tcstem.jsonl gpurun_out/exp_tcstem.err
timeout 150 python -m pytest tests/test_gpu_model.py -q -m gpu -k "tensor_core_stem" -s --timeout 120 --timeout-method=thread -x > gpurun_out/pytest_tcstem.log 2>&1; rc=$?; echo "tcstem test rc=$rc"; grep -E "tc stem|passed|failed|Error|error" gpurun_out/pytest_tcstem.log | head -20
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_tcstem.log; exit 0; fi
OGL_FUSE_STEM=2 timeout 600 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu_mode2.log 2>&1; echo "pytest (OGL_FUSE_STEM=2) rc=$?"; tail -6 gpurun_out/pytest_gpu_mode2.log
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_tcstem.jsonl 2>> gpurun_out/exp_tcstem.err; }
run OGL_FUSE_STEM=1
run OGL_FUSE_STEM=2
run OGL_FUSE_STEM=1
run OGL_FUSE_STEM=2
run OGL_FUSE_STEM=2 OGL_EXPERIMENT=1 OGL_DBG=64
python scripts/show_exp.py gpurun_out/exp_tcstem.jsonl | cut -c1-120; tail -3 gpurun_out/exp_tcstem.err
