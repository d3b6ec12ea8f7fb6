mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_layers.py -q -m gpu --timeout 90 --timeout-method=thread -s > gpurun_out/layers.log 2>&1; echo "layers rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_features.py -q -m gpu --timeout 120 --timeout-method=thread -s > gpurun_out/features.log 2>&1; echo "features rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 --timeout-method=thread -s > gpurun_out/model.log 2>&1; echo "model rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -30 gpurun_out/layers.log
