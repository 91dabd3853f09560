mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -x 2>&1 | grep -v "UserWarning\|kaiming\|WeightNorm" > gpurun_out/r2f_train.log; tail -40 gpurun_out/r2f_train.log
timeout 1800 python -m pytest tests -q -m gpu -s --deselect tests/test_gpu_train.py 2>&1 | grep -v "UserWarning\|kaiming\|WeightNorm" > gpurun_out/r2f_gpu_suite.log
grep -n "passed\|failed\|FAILED" gpurun_out/r2f_gpu_suite.log | tail -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 1500 gpurun_out/r2f_bench.json
