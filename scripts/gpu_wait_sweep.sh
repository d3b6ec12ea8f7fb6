# usage: gpu_wait_sweep.sh  -- GPU tests, then A/B of the waiting schemes (OGL_WAIT_*) and, when
# openglottal_b200/lib/exp/lib_v9.so exists, of the previous build (OGL_LIB), all on one box
mkdir -p gpurun_out
rm -f gpurun_out/exp_wait.jsonl gpurun_out/exp_wait.err
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" gpurun_out/pytest_gpu.log | head -20; exit 0; fi
run() { env "$@" timeout 120 python scripts/layer_times.py 512 4 "$*" >> gpurun_out/exp_wait.jsonl 2>> gpurun_out/exp_wait.err; }
OLD=openglottal_b200/lib/exp/lib_v9.so
[ -f $OLD ] && run OGL_LIB=$OLD
run OGL_X=base
run OGL_WAIT_SLEEP=256
run OGL_WAIT_SLEEP=1000
run OGL_WAIT_SLEEP=0 OGL_WAIT_HINT=1000
run OGL_WAIT_SLEEP=0 OGL_WAIT_HINT=20000
run OGL_WAIT_SLEEP=0 OGL_WAIT_HINT=1000 OGL_WAIT_HINT_CRIT=1000
run OGL_WAIT_SLEEP=64 OGL_WAIT_HINT_CRIT=1000
run OGL_WAIT_STEM_RELAXED=1
run OGL_WAIT_STEM_RELAXED=1 OGL_WAIT_SLEEP=256
[ -f $OLD ] && run OGL_LIB=$OLD
run OGL_X=base
python scripts/show_exp.py gpurun_out/exp_wait.jsonl; tail -3 gpurun_out/exp_wait.err
