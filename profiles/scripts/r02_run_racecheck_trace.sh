# one B200: racecheck again with the hazard list kept, and the in-kernel clock64 timeline of the slot kernel
mkdir -p gpurun_out
SEL='two-hop-3-64-2 or two-hop-3-128-2 or gather+reduce-3-32-2 or test_node_linear_tensor_core_path[True-64-64] or test_node_linear_wgrad[True-128-128] or test_halo or test_rank_topk_vs_oracle[192 or fused_adam_bit_identical_with_torch[0.0]'
timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > /tmp/race.log 2>&1
grep -E "^========= (Error|Warning|RACECHECK)|hazards\]" /tmp/race.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -60 > gpurun_out/r2_racecheck_hazards.txt
grep -E "=========     (Write|Read) Thread|=========     Current Value" /tmp/race.log | head -5 >> gpurun_out/r2_racecheck_hazards.txt
grep -B2 -A12 "^========= Error" /tmp/race.log | head -120 > gpurun_out/r2_racecheck_examples.txt
python profiles/trace_slot_kernel.py cikm > gpurun_out/r2_trace_slot_cikm.txt 2>&1
python profiles/trace_slot_kernel.py amazon-full > gpurun_out/r2_trace_slot_amazon-full.txt 2>&1
