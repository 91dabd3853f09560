mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -s -x 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" > gpurun_out/r2m_train.log; grep -n "wgrad\|passed\|failed\|FAILED\|Error\|error" gpurun_out/r2m_train.log | head -30
timeout 600 python bench.py --config train --steps 10 --warmup 5 2>/dev/null | cut -c1-220
timeout 600 python - <<'PY' 2>/dev/null | cut -c1-220
import sys, subprocess
sys.argv=['bench.py','--config','train','--steps','10','--warmup','5']
import hold_b200.train as T
T.WGRAD_TC=False
import runpy; runpy.run_path('bench.py', run_name='__main__')
PY
