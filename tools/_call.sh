mkdir -p gpurun_out
timeout 600 python tools/prof_train.py 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" > gpurun_out/r2k_prof_train.log; grep -n "plain step\|^step\|Self CUDA time" gpurun_out/r2k_prof_train.log; grep -n "hold::\|cutlass\|indexing\|Backward" gpurun_out/r2k_prof_train.log | head -30 | cut -c1-220
timeout 900 python tools/bench_aux.py 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tee gpurun_out/r2k_bench_aux.log
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_stages.py -q -x 2>&1 | tail -3
