# final validation of the round-2 build: GPU suite and smoke (the default bench line of this build:
# profiles/bench_r02_v4.json, taken by the previous call of this script)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread -x -s > gpurun_out/r2_pytest_v4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_v4.log
grep -h "gaw 512x256" gpurun_out/r2_pytest_v4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_v4.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke_v4.log
