# two B200s: cheap validation of the workloads the N = 8 session runs
mkdir -p gpurun_out
run() { name=$1; shift; envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 "$@" > gpurun_out/r2_chk_$name.json 2> gpurun_out/r2_chk_$name.err
  echo "$name rc=$? $(head -c 400 gpurun_out/r2_chk_$name.json)"; tail -3 gpurun_out/r2_chk_$name.err; }
(timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -4) > gpurun_out/r2_chk_tests.log
run scaled IHG_OVERLAP=0 --workload scaled --scale 0.1
run rank IHG_OVERLAP=0 --workload rank
run amazon IHG_OVERLAP=0 --workload amazon-full
run amazon_strong IHG_OVERLAP=0 --workload amazon-full --scaling strong
