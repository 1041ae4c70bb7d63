# one B200: quick GPU tests of the new pieces, the default bench line, the ranking workload, the CPU arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "adam or graphed" 2>&1 | tail -5 > gpurun_out/r2_t2_tests.log
python bench.py > gpurun_out/r2_b1_default.json 2> gpurun_out/r2_b1_default.err
python bench.py --workload rank > gpurun_out/r2_b1_rank.json 2> gpurun_out/r2_b1_rank.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_b1_ref.json 2> gpurun_out/r2_b1_ref.err
