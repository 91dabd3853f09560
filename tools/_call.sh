mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -q -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -25
