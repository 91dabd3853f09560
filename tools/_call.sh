mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s 2>&1 | grep -v -i "warn\|kaiming\|WeightNorm" > gpurun_out/r2d_gpu_suite.log
grep -n "passed\|failed\|FAILED" gpurun_out/r2d_gpu_suite.log | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -c 2 -f -o gpurun_out/r2d_tc python tools/prof_kernels.py > gpurun_out/r2d_ncu_tc.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 2500 gpurun_out/r2d_bench.json
