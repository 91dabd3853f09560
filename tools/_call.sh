mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep "smoke OK\|rror"
timeout 1200 python bench.py 2>gpurun_out/r3o_bench.err | tee gpurun_out/r3o_bench.json | cut -c1-200
