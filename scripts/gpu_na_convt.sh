mkdir -p gpurun_out
rm -f gpurun_out/exp_na.jsonl gpurun_out/exp_na.err
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_na.jsonl 2>> gpurun_out/exp_na.err; }
run OGL_NA_CONVT=3
run OGL_NA_CONVT=4
run OGL_NA_CONVT=5
run OGL_NA_CONVT=6
run OGL_NA_CONVT=3
run OGL_NA_CONVT=5
python scripts/show_exp.py gpurun_out/exp_na.jsonl | cut -c1-230; tail -3 gpurun_out/exp_na.err
