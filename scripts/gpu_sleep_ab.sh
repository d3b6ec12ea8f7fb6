mkdir -p gpurun_out
rm -f gpurun_out/exp_sl.jsonl gpurun_out/exp_sl.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_sl.jsonl 2>> gpurun_out/exp_sl.err; }
run OGL_WAIT_SLEEP=64
run OGL_WAIT_SLEEP=0
run OGL_WAIT_SLEEP=20
run OGL_WAIT_SLEEP=64
run OGL_WAIT_SLEEP=0
python scripts/show_exp.py gpurun_out/exp_sl.jsonl | cut -c1-230; tail -3 gpurun_out/exp_sl.err
