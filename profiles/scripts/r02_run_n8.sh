# eight B200s: multi-GPU parity tests (4 ranks), weak scaling at N = 8 for both conv workloads (serial and
# overlapped exchange), the 100 M-hyperedge workload (BASELINE.json configs[3]) and the ranking workload
mkdir -p gpurun_out
T=${TAG:-r2_n8}
run() { # name, env, args...
  name=$1; shift; envs=$1; shift
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  echo "$name rc=$? $(head -c 300 gpurun_out/${T}_$name.json)"
}
(timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -6) > gpurun_out/${T}_tests.log
run amazon-full IHG_OVERLAP=0 --workload amazon-full
run cikm IHG_OVERLAP=0 --workload cikm
# (the *_overlap lines of this session came from commit 1bd0c3d's opt-in IHG_OVERLAP=1 path, since removed)
run scaled IHG_OVERLAP=0 --workload scaled --steps 10 --warmup 3
run rank IHG_OVERLAP=0 --workload rank
nvidia-smi --query-gpu=index,memory.used --format=csv > gpurun_out/${T}_mem.log
