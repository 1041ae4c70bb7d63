# A/B of a hoisted form of the tensor-core forward (first-order blocks as a node-level typed Linear whose rows
# the kernel's epilogue adds) against the shipped un-hoisted form.  The hoisted variant and bench.py's
# --hoist-min-dim switch lived only in the working tree of this experiment (results: profiles/r02_ab_fwd_hoist_*;
# hoisted was slower: cikm 6.42 vs 4.26 ms, amazon-full 0.65 vs 0.38 ms) and were removed again; see DESIGN.md 9.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "interactor_forward_forms or medium_graph or model_forward_backward or implementations_agree" 2>&1 | tail -8 > gpurun_out/r3_tests.log
for w in cikm amazon-full; do
  for h in 0 1000; do
    python bench.py --workload $w --also none --steps 10 --warmup 3 --no-cpu-baseline --hoist-min-dim $h > gpurun_out/r3_bench_${w}_hoist$h.json 2> gpurun_out/r3_bench_${w}_hoist$h.err
  done
done
