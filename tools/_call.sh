mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -s 2>&1 | grep -v "Warning\|kaiming\|WeightNorm" > gpurun_out/r3n_train.log; grep -n "eager losses\|passed\|failed\|FAILED\|^E " gpurun_out/r3n_train.log | head
timeout 600 python bench.py --config train --steps 20 --warmup 5 --graph 2>gpurun_out/r3n_train_graph.err | tee gpurun_out/r3n_train_graph.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('graph run:', d.get('ms_per_step'), d.get('launch_mode'), d.get('loss'), d.get('unavailable'), d.get('gpu_launches'))"
timeout 600 python bench.py --config train --steps 20 --warmup 5 2>gpurun_out/r3n_train_eager.err | tee gpurun_out/r3n_train_eager.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('eager run:', d.get('ms_per_step'), d.get('launch_mode'), d.get('loss'), d.get('gpu_launches'))"
grep -n "Error" gpurun_out/r3n_train_graph.err | head -3
