mkdir -p gpurun_out
python scripts/profile_forward.py 512 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_forward.py 512 > gpurun_out/ncu1.log 2>&1
echo "launches rc=$?"
python scripts/profile_forward.py 128 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 21 -c 21 -o gpurun_out/prof_conv_tc -f python scripts/profile_forward.py 128 > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/; tail -3 gpurun_out/ncu2.log
