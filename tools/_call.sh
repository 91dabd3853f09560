mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_background.py -q -s -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" > gpurun_out/r2o_train.log; grep -n "background\|passed\|failed\|FAILED\|Error\|error\|losses" gpurun_out/r2o_train.log | head -60
timeout 600 python bench.py --config train --steps 10 --warmup 5 2>gpurun_out/r2o_train_bench.err | tee gpurun_out/r2o_train_bench.json | cut -c1-260
tail -5 gpurun_out/r2o_train_bench.err
