mkdir -p gpurun_out
timeout 900 python tools/exp_r2b.py 2>&1 | grep -v -i "warn\|kaiming\|WeightNorm" | tee gpurun_out/r2b_exp.log
timeout 1500 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -v -i "warn\|kaiming\|WeightNorm" | tail -150 > gpurun_out/r2b_gpu_suite.log
tail -30 gpurun_out/r2b_gpu_suite.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.json
