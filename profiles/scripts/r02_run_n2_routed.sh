# two B200s: routed reduce (reduction kernels write the halo partial sums straight into the owners' receive buffers)
# against the pull path (IHG_ROUTED_REDUCE=0): parity tests, then the cikm / amazon-full weak-scaling lines
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -x -q -k "sharded or routed" 2>&1 | tail -6) > gpurun_out/r5_n2_tests.log
run() { name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/r5_n2_$name.json 2> gpurun_out/r5_n2_$name.err
  echo "$name rc=$? $(head -c 300 gpurun_out/r5_n2_$name.json)"; tail -3 gpurun_out/r5_n2_$name.err; }
run cikm_routed IHG_ROUTED_REDUCE=1 --workload cikm
run cikm_pull IHG_ROUTED_REDUCE=0 --workload cikm
run amazon_routed IHG_ROUTED_REDUCE=1 --workload amazon-full
run amazon_pull IHG_ROUTED_REDUCE=0 --workload amazon-full
