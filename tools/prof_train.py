"""Where a training step's time goes: host-side cProfile of TrainStep.step and a torch.profiler kernel table."""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hold_b200 import capi, scene_io, synth, train

ctx = capi.Context(0); dev = torch.device("cuda", 0)
Bf, px = 10, 128
sc = synth.make_scene(H=bench.H, W=bench.W, S=bench.S, nodes=bench.NODES, B=Bf, seed=0)
for nid in sc.node_ids:
    sc.beta[nid] = torch.tensor(bench.BETA)
net = scene_io.build_net(sc, ctx, capi.MLP_TC)
gen = torch.Generator().manual_seed(100)
ids = torch.stack([torch.randperm(bench.H * bench.W, generator=gen)[:px] for _ in range(Bf)])
inp = scene_io.scene_input(sc, dev)
inp["uv"] = torch.gather(inp["uv"], 1, ids.to(dev)[:, :, None].expand(-1, -1, 2)).contiguous()
R = Bf * px
gt_rgb = torch.rand(R, 3, device=dev); gt_mask = torch.zeros(R, 4, device=dev); gt_mask[:, 0] = 1
ts = train.TrainStep(net)
for _ in range(2):
    ts.step(inp, gt_rgb, gt_mask)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    ts.step(inp, gt_rgb, gt_mask)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    ts.step(inp, gt_rgb, gt_mask)
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
