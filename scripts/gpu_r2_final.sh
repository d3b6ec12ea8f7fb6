# Final evidence of round 2 (one GPU): smoke(), the default bench line (with the cpu_baseline leg) and
# the reference arm as the driver runs them, then ncu: a full capture of the 17 tcgen05 launches of
# one forward (batch 128) and the launch list of a short bench.py run. Each ncu run follows a plain
# run of the same command that exited 0.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py --layers-out gpurun_out/r2_layers_final.json > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_final.json', 'gpurun_out/r2_bench_final_reference.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 3), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
PY
python scripts/profile_forward.py 128 > gpurun_out/r2_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:conv_tc_kernel|s2d_tc_kernel|upcat_tc_kernel' -s 17 -c 17 -o gpurun_out/prof_tc_r02_v3 -f python scripts/profile_forward.py 128 > gpurun_out/r2_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu_full.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "launches rc=$?"
