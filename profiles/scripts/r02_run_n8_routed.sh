# eight B200s: cikm weak scaling, routed reduce (default) vs pull
mkdir -p gpurun_out
run() { name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 10 --warmup 3 "$@" > gpurun_out/r5_n8_$name.json 2> gpurun_out/r5_n8_$name.err
  echo "$name rc=$? $(head -c 260 gpurun_out/r5_n8_$name.json)"; tail -2 gpurun_out/r5_n8_$name.err; }
run cikm_routed IHG_ROUTED_REDUCE=1 --workload cikm
run cikm_pull IHG_ROUTED_REDUCE=0 --workload cikm
