mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -3
for n in 0 12 0 12; do timeout 300 python tools/exp_epilogue.py --run $n 2>&1 | grep "^exp"; done | tee gpurun_out/r3j_exp_unroll.log
