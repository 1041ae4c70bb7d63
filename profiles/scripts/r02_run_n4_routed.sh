# four B200s: multi-peer routes (world 4 parity tests), cikm weak scaling routed vs pull
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "(sharded_stack and routed-reduce) or training_step" 2>&1 | tail -4) > gpurun_out/r5_n4_tests.log
run() { name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 "$@" > gpurun_out/r5_n4_$name.json 2> gpurun_out/r5_n4_$name.err
  echo "$name rc=$? $(head -c 260 gpurun_out/r5_n4_$name.json)"; tail -2 gpurun_out/r5_n4_$name.err; }
run cikm_routed IHG_ROUTED_REDUCE=1 --workload cikm
run cikm_pull IHG_ROUTED_REDUCE=0 --workload cikm
