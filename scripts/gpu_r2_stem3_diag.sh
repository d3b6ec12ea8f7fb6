# Fused-stem launch: sensitivity to the hand-off latency of the mbarrier waits
mkdir -p gpurun_out
: > gpurun_out/r2_exp_stem3diag4.jsonl
for cfg in "OGL_WAIT_SLEEP=64" "OGL_WAIT_SLEEP=0" "OGL_WAIT_SLEEP=20" "OGL_WAIT_SLEEP=0 OGL_EXPERIMENT=1 OGL_DBG=1029" "OGL_WAIT_SLEEP=0 OGL_FUSE_STEM=2" "OGL_WAIT_HINT=200 OGL_WAIT_SLEEP=0"; do
  env $cfg timeout 200 python scripts/layer_times.py 512 4 "$cfg" >> gpurun_out/r2_exp_stem3diag4.jsonl 2>> gpurun_out/r2_exp_stem3diag4.err
done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_stem3diag4.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in L.items() if 'downs.0' in k or 'head' in k}, 'Mcycles', {k: round(v * d['clocks']['sm_mhz'] / 1e3, 3) for k, v in L.items() if 'downs.0' in k or 'head' in k})
PY
timeout 300 python -m pytest tests/test_gpu_pipelines.py tests/test_gpu_cli.py -m gpu -x -q > gpurun_out/r2_decode_tests.log 2>&1; echo "pipeline tests rc=$?"; tail -3 gpurun_out/r2_decode_tests.log
timeout 600 python scripts/ingest_bench.py 40000 MJPG > gpurun_out/r2_ingest_mjpg.json 2> gpurun_out/r2_ingest_mjpg.err; echo "ingest rc=$?"; cat gpurun_out/r2_ingest_mjpg.json; tail -3 gpurun_out/r2_ingest_mjpg.err
