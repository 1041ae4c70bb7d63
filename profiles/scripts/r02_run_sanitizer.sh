# one B200: compute-sanitizer memcheck + racecheck + synccheck over a subset of the GPU parity tests that
# runs every warp-specialised tcgen05 / mbarrier kernel (forward, slot gradients, dW_hi, typed Linear and its
# weight gradient) plus the reductions, ranking, Adam and the halo kernels
mkdir -p gpurun_out
SEL='two-hop-3-64-2 or two-hop-3-128-2 or gather+reduce-3-32-2 or test_node_linear_tensor_core_path[True-64-64] or test_node_linear_wgrad[True-128-128] or test_halo or test_rank_topk_vs_oracle[192 or fused_adam_bit_identical_with_torch[0.0]'
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/r2_sanitizer_$tool.log | tail -5 >> gpurun_out/r2_sanitizer_summary.txt
  tail -c 4000 gpurun_out/r2_sanitizer_$tool.log > gpurun_out/r2_sanitizer_$tool.tail.log; rm gpurun_out/r2_sanitizer_$tool.log
done
cat gpurun_out/r2_sanitizer_summary.txt
