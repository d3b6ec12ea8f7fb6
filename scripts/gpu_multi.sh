mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo "long rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
for f in ('bench_long','bench_n2','bench_ref_n2'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'clocks', d.get('clocks'), 'n', d['n_gpus'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_n2.err
