mkdir -p gpurun_out
python scripts/profile_forward.py 128 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 21 -c 21 -o gpurun_out/prof_conv_tc_v3 -f python scripts/profile_forward.py 128 > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu2.log
