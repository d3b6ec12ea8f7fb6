# usage: gpu_quick.sh <tag> [ENV=VAL ...]  -- GPU tests, then per-launch times of the default schedule
mkdir -p gpurun_out
tag=$1; shift
rm -f gpurun_out/exp_$tag.jsonl gpurun_out/exp_$tag.err
timeout 600 python -m pytest tests -q -m gpu --timeout 200 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu.log | head -20; exit 0; fi
run() { env "$@" timeout 120 python scripts/layer_times.py 512 5 "$*" >> gpurun_out/exp_$tag.jsonl 2>> gpurun_out/exp_$tag.err; }
run OGL_S2D=1
for e in "$@"; do run $e; done
python scripts/show_exp.py gpurun_out/exp_$tag.jsonl; tail -3 gpurun_out/exp_$tag.err
