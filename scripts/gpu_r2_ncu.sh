# ncu evidence for round 2 (one GPU): full capture of the 17 tcgen05 launches of one forward
# (batch 128), then the launch list of a short bench.py run. Each ncu run follows a plain run of
# the same command that exited 0.
mkdir -p gpurun_out
python scripts/profile_forward.py 128 > gpurun_out/r2_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:conv_tc_kernel|s2d_tc_kernel|upcat_tc_kernel' -s 17 -c 17 -o gpurun_out/prof_tc_r02_v1 -f python scripts/profile_forward.py 128 > gpurun_out/r2_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu_full.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
echo "launches rc=$?"
