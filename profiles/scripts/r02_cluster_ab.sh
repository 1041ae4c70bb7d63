# dev A/B, working tree only (not shipped): (1) IHG_DEV_CLUSTER switch in the forward / slot-gradient launchers = cluster size
# of the weight multicast; (2) slot kernel reading its def rows straight from global memory (thread = row) instead of
# the cp.async staging, the freed 32 KB used as a fourth weight stage.  Results: profiles/r02_ab_slot_directdef_*:
# cluster 1 / 2 / 4 within 2 % of each other; (2) slower (edge_interact_bwd 7.49 vs 6.96 ms at cikm, 0.705 vs 0.64 ms at
# amazon-full).  Both removed again; see DESIGN.md section 9.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "graph2d or gcn or medium_graph or model_forward_backward or conv_stack" 2>&1 | tail -5 > gpurun_out/r4_tests.log
for w in cikm amazon-full; do
  for c in 1 2 4; do
    IHG_DEV_CLUSTER=$c python bench.py --workload $w --also none --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4_bench_${w}_cl$c.json 2> gpurun_out/r4_bench_${w}_cl$c.err
  done
done
