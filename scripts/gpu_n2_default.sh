mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 > gpurun_out/bench_n2_default.json 2> gpurun_out/bench_n2_default.err; echo "N=2 default rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_default.json').read().strip().splitlines()[-1])
print(d['steps'], d['warmup'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref N=2 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n2.json
