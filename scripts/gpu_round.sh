# Round evidence: bench line (with CPU baseline), ncu launch list of bench.py, full ncu capture
# of every tcgen05 launch of one forward (batch 128).
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'roofline frac', round(d['roofline']['frac'],3), 'clocks', d['clocks'], 'cpu', d.get('cpu_baseline'))
for l in json.load(open('gpurun_out/layers_n1.json')): print('%-32s %8.3f ms %8.1f TF' % (l['layer'], l['ms'], l['tflops'] or 0))
PY
tail -3 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "launches rc=$?"
python scripts/profile_forward.py 128 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:conv_tc_kernel|s2d_tc_kernel' -s 20 -c 20 -o gpurun_out/prof_tc_r01_v12 -f python scripts/profile_forward.py 128 > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu2.log
