mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp1.jsonl 2>> gpurun_out/exp1.err; }
rm -f gpurun_out/exp1.jsonl gpurun_out/exp1.err
run OGL_DBG=0
run OGL_DBG=2
run OGL_DBG=4
run OGL_DBG=1
run OGL_DBG=5
run OGL_DBG=8
run OGL_DBG=16
run OGL_DBG=24
run OGL_DBG=25
run OGL_DBG=0 OGL_NA=2
run OGL_DBG=0 OGL_NA=4
run OGL_DBG=0 OGL_NW=3
wc -l gpurun_out/exp1.jsonl; tail -3 gpurun_out/exp1.err
