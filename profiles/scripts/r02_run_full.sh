# one B200: the whole GPU test-suite, smoke, the default bench line
mkdir -p gpurun_out
T=${TAG:-r2_full}
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${T}_tests.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1 > gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
