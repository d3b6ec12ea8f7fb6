mkdir -p gpurun_out
for m in 2 1; do
OGL_FUSE_STEM=$m timeout 300 ncu --set full --clock-control none --import-source on -k regex:s2d_tc_kernel -s 3 -c 1 -o gpurun_out/prof_stem_mode$m -f python scripts/profile_forward.py 128 > gpurun_out/ncu_stem$m.log 2>&1; echo "mode $m rc=$?"; tail -2 gpurun_out/ncu_stem$m.log
done
