mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sampler_rounds.py tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_edges.py -q -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -3
timeout 900 python bench.py --no-extras 2>gpurun_out/r3f_bench.err | tee gpurun_out/r3f_bench.json | cut -c1-260
