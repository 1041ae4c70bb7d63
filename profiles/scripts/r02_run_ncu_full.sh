# one B200: ncu --set full of the three tensor-core interaction kernels of one eager cikm conv step
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k regex:"interact_bwd_slot_ts_kernel|edge_interact_bwd_wgrad_tc_kernel|feature_interact_fwd_ts_kernel" \
  -o /tmp/ncu_full_cikm -f python bench.py --workload cikm --also none --steps 3 --warmup 3 --no-cpu-baseline --no-graph --profile-step > gpurun_out/r2f_ncu_full_cikm.log 2>&1
ncu -i /tmp/ncu_full_cikm.ncu-rep --page details > gpurun_out/r2f_ncu_full_cikm.txt 2>> gpurun_out/r2f_ncu_full_cikm.log
ls -la /tmp/ncu_full_cikm.ncu-rep >> gpurun_out/r2f_ncu_full_cikm.log
