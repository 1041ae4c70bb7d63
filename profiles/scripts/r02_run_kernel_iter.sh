# one B200: parity of the interaction kernels, slot-kernel timeline, default bench line
mkdir -p gpurun_out
T=${TAG:-r2_k1}
(python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "medium_graph or model_forward or conv_stack or stock_torch or conv_fwd_bwd_vs_oracle or interactor" 2>&1 | tail -4) > gpurun_out/${T}_tests.log
python profiles/trace_slot_kernel.py cikm > gpurun_out/${T}_trace_slot_cikm.txt 2>&1
python profiles/trace_slot_kernel.py amazon-full > gpurun_out/${T}_trace_slot_amazon-full.txt 2>&1
python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
