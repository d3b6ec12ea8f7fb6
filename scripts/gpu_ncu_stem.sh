# full ncu capture of the fused-stem kernel in modes 2 (tensor cores) and 1 (CUDA cores), then
# the energy of that launch in both modes
mkdir -p gpurun_out
for m in 2 1; do
OGL_FUSE_STEM=$m timeout 300 ncu --set full --clock-control none --import-source on -k regex:s2d_tc_kernel -s 3 -c 1 -o gpurun_out/prof_stem_mode$m -f python scripts/profile_forward.py 128 > gpurun_out/ncu_stem$m.log 2>&1; echo "mode $m rc=$?"; tail -2 gpurun_out/ncu_stem$m.log
done
for m in 1 2 1 2; do
OGL_FUSE_STEM=$m timeout 200 python scripts/layer_energy.py 512 1.5 9 0 > gpurun_out/energy_stem$m.json 2>> gpurun_out/energy_stem.err; echo "energy mode $m rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/energy_stem$m.json')); r=d['layers'][0]; print('mode $m', r['layer'], round(r['joule'],3),'J', round(r['ms'],3),'ms  forward', round(d['forward']['joule'],3),'J', round(d['forward']['ms'],3),'ms')"
done
