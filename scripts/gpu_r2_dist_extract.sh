# file -> features on 1, 2, 4 ... GPUs of one box (gpurun --gpus N): sharded decode + segmentation
mkdir -p gpurun_out
: > gpurun_out/r2_dist_extract.jsonl
python scripts/dist_extract_check.py 40000 2> gpurun_out/r2_dist_extract_n1.err | tail -1 | tee -a gpurun_out/r2_dist_extract.jsonl
for n in ${1:-2}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 scripts/dist_extract_check.py 40000 2> gpurun_out/r2_dist_extract_n$n.err | tail -1 | tee -a gpurun_out/r2_dist_extract.jsonl
  echo "N=$n rc=$?"; tail -2 gpurun_out/r2_dist_extract_n$n.err
done
nproc
