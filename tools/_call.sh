mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_e2e.py -q -s 2>&1 | grep -v "UserWarning\|kaiming\|WeightNorm" > gpurun_out/r2g_train.log; grep -n "grad \|passed\|failed\|FAILED\|Error" gpurun_out/r2g_train.log | tail -70
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -15
timeout 600 python bench.py --config train --steps 5 --warmup 2 > gpurun_out/r2g_train_bench.json 2> gpurun_out/r2g_train_bench.err; tail -c 1500 gpurun_out/r2g_train_bench.json; tail -5 gpurun_out/r2g_train_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 2 -c 1 -f -o gpurun_out/r2g_tc0_bench python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2g_ncu_tc0.log 2>&1; tail -2 gpurun_out/r2g_ncu_tc0.log | cut -c1-300
