# Where the fused-stem launch spends its time: parts switched off (results are garbage with OGL_DBG).
mkdir -p gpurun_out
: > gpurun_out/r2_exp_stem3diag2.jsonl
for cfg in "OGL_DBG=0" "OGL_DBG=256" "OGL_DBG=512" "OGL_DBG=768" "OGL_DBG=769" "OGL_STEM_LO=0 OGL_DBG=768"; do
  env $cfg timeout 200 python scripts/layer_times.py 512 4 "$cfg" >> gpurun_out/r2_exp_stem3diag2.jsonl 2>> gpurun_out/r2_exp_stem3diag2.err
done
python - <<'PY'
import json
for line in open('gpurun_out/r2_exp_stem3diag2.jsonl'):
    d = json.loads(line)
    L = d['layers']
    print(d['tag'], 'step', round(d['ms_step'], 3), 'sm', d['clocks']['sm_mhz'], {k: round(v, 4) for k, v in L.items() if 'downs.0' in k or 'head' in k})
PY
timeout 600 python scripts/ingest_bench.py 20000 MJPG > gpurun_out/r2_ingest_mjpg.json 2> gpurun_out/r2_ingest_mjpg.err; echo "ingest rc=$?"; cat gpurun_out/r2_ingest_mjpg.json; tail -3 gpurun_out/r2_ingest_mjpg.err
timeout 600 python scripts/ingest_bench.py 8000 FFV1 > gpurun_out/r2_ingest_ffv1.json 2> gpurun_out/r2_ingest_ffv1.err; echo "ingest rc=$?"; cat gpurun_out/r2_ingest_ffv1.json; tail -3 gpurun_out/r2_ingest_ffv1.err
nproc; ls /usr/lib/x86_64-linux-gnu | grep -i "cuvid\|nvidia-encode" | head
