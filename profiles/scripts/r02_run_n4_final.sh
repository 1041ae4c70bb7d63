# four B200s, final code: multi-GPU parity tests (world 4) and the amazon-full weak-scaling line
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -4) > gpurun_out/r6_n4_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 4 --steps 10 --warmup 3 --workload amazon-full > gpurun_out/r6_n4_amazon-full.json 2> gpurun_out/r6_n4_amazon-full.err
echo "rc=$? $(head -c 260 gpurun_out/r6_n4_amazon-full.json)"
