mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -4
timeout 900 python bench.py --no-extras 2>gpurun_out/r2y_bench.err | tee gpurun_out/r2y_bench.json | cut -c1-300
