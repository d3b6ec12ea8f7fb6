# byte / integer operators: new build against the previous one in the same process (bit-equality,
# GB/s against the HBM roof), then the whole GPU suite and smoke() on the new build
mkdir -p gpurun_out
timeout 150 python scripts/byte_ops_bench.py --prev openglottal_b200/lib/prev_r02_v5.so --out gpurun_out/byte_ops_r02_v6.json > gpurun_out/byte_ops_r02_v6.log 2>&1; echo "byte_ops rc=$?"; tail -1 gpurun_out/byte_ops_r02_v6.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/byte_ops_r02_v6.json"))
    for r in d["rows"]:
        if "ms" in r:
            print(f'{r["op"][:52]:52s} {r["ms"]:7.3f} ms {r["gbs"]:7.0f} GB/s {r["frac_of_hbm_peak"]:.2f} | prev {r.get("prev_ms", 0):7.3f} ms eq={r.get("equal_to_prev_build")}')
        else:
            print(f'{r["op"][:70]:70s} eq={r.get("equal_to_prev_build")} {r.get("error", "")}')
except Exception as e:
    print("no json", e)
PY
timeout 400 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread -x -s > gpurun_out/r2_pytest_v6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_v6.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_v6.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_v6.log
