"""Time one SDF-net launch of a sampler round (262144 x 128 points) — the roofline kernel of bench.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import capi, scene_io, synth
ctx = capi.Context(0); dev = torch.device("cuda", 0)
sc = synth.make_scene(H=8, W=8, S=128, nodes=("right", "object"))
net = scene_io.build_net(sc, ctx, capi.MLP_TC)
P = 262144 * 128
xc = (torch.rand(P, 3, device=dev) - 0.5) * 1.6
sdf = torch.empty(P, device=dev)
node = net.nodes["right"]
def launch():
    capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(xc), None, capi.ptr(sdf), None, None, capi.stream_ptr()))
for _ in range(2): launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): launch()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"HOLD_TC_DBG={os.environ.get('HOLD_TC_DBG','0')}: {ms:.2f} ms per launch, {2*459008*P/ms/1e9:.1f} TFLOP/s algorithmic")
