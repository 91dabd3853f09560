mkdir -p gpurun_out
for n in 101 0 101 0; do timeout 300 python tools/exp_epilogue.py --run $n 2>&1 | grep "^exp"; done | tee gpurun_out/r3g_exp_rev.log
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_edges.py -q -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -3
