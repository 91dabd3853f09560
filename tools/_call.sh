mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" > gpurun_out/r3c_train.log; grep -n "forward_train\|passed\|failed\|FAILED\|Error\|error\|losses\|loss/" gpurun_out/r3c_train.log | head -40; tail -30 gpurun_out/r3c_train.log | grep -v "^tests/"
