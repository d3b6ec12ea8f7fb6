mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py -k "cta_pairs" -q -m gpu --timeout 60 --timeout-method=thread -x -s > gpurun_out/pair_layers.log 2>&1; rc=$?; echo "pair layers rc=$rc"; tail -25 gpurun_out/pair_layers.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python -m pytest tests/test_gpu_model.py -k "cta_pairs" -q -m gpu --timeout 100 --timeout-method=thread -x -s > gpurun_out/pair_model.log 2>&1; rc=$?; echo "pair model rc=$rc"; tail -8 gpurun_out/pair_model.log
if [ $rc -ne 0 ]; then exit 0; fi
rm -f gpurun_out/exp_pair.jsonl gpurun_out/exp_pair.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp_pair.jsonl 2>> gpurun_out/exp_pair.err; }
run OGL_CG=1
run OGL_CG=2
run OGL_CG=1
run OGL_CG=2
python scripts/show_exp.py gpurun_out/exp_pair.jsonl; tail -3 gpurun_out/exp_pair.err
