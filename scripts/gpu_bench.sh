mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" > gpurun_out/rc.txt
cat gpurun_out/rc.txt; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; cat gpurun_out/layers_n1.json | python -c "
import json,sys
for l in json.load(sys.stdin): print('%-24s %8.3f ms %8.1f TF' % (l['layer'], l['ms'], l['tflops'] or 0))"
