# two B200s: the multi-GPU parity tests, then the sharded bench (weak scaling, N = 2)
mkdir -p gpurun_out
T=${TAG:-r2_n2}
(timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15) > gpurun_out/${T}_tests.log
for w in ${WORKLOADS:-amazon-full cikm}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --workload $w ${BENCH_ARGS} > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
  tail -c 600 gpurun_out/${T}_$w.json
done
