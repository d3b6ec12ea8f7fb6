# transposed conv on CTA pairs (OGL_CONVT_PAIR=1): layer tests under a short timeout, the whole GPU
# suite, then A/B on one box
mkdir -p gpurun_out
rm -f gpurun_out/exp_ctp.jsonl gpurun_out/exp_ctp.err
OGL_CONVT_PAIR=1 timeout 150 python -m pytest tests/test_gpu_layers.py -q -m gpu -k "convt" --timeout 100 --timeout-method=thread -x > gpurun_out/pytest_ctp.log 2>&1; rc=$?; echo "convT pair layer tests rc=$rc"; tail -3 gpurun_out/pytest_ctp.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_ctp.log | head; exit 0; fi
OGL_CONVT_PAIR=1 timeout 600 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu_ctp.log 2>&1; echo "pytest (OGL_CONVT_PAIR=1) rc=$?"; tail -4 gpurun_out/pytest_gpu_ctp.log
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_ctp.jsonl 2>> gpurun_out/exp_ctp.err; }
run OGL_CONVT_PAIR=0
run OGL_CONVT_PAIR=1
run OGL_CONVT_PAIR=0
run OGL_CONVT_PAIR=1
python scripts/show_exp.py gpurun_out/exp_ctp.jsonl | cut -c1-230; tail -3 gpurun_out/exp_ctp.err
