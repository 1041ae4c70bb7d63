# one B200, end of round 2: the whole GPU test-suite, smoke, the default bench line (+ rank, reference arm), then the
# ncu per-kernel metrics / launch lists of one eager step per workload
mkdir -p gpurun_out
T=r2f
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${T}_tests.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1 > gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --workload rank > gpurun_out/${T}_rank.json 2> gpurun_out/${T}_rank.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
python bench.py --workload amazon-small --also none > gpurun_out/${T}_amazon-small.json 2> gpurun_out/${T}_amazon-small.err
TAG=r2f bash profiles/scripts/r02_run_ncu.sh > gpurun_out/${T}_ncu.log 2>&1
