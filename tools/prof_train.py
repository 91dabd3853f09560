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
pose_leaves = []
if os.environ.get("POSE_GRAD", "1") == "1":
    for k in list(inp):
        if torch.is_tensor(inp[k]) and inp[k].is_floating_point() and any(k.endswith(sfx) for sfx in (".full_pose", ".transl", ".global_orient")):
            inp[k] = inp[k].clone().requires_grad_(True)
            pose_leaves.append(inp[k])
ts = train.TrainStep(net)
ts.params = ts.params + pose_leaves
for _ in range(2):
    ts.step(inp, gt_rgb, gt_mask)
torch.cuda.synchronize()
import time
for i in range(3):
    t0 = time.perf_counter(); ts.step(inp, gt_rgb, gt_mask); torch.cuda.synchronize()
    print(f"plain step {i}: {1e3 * (time.perf_counter() - t0):.1f} ms (pose leaves: {len(pose_leaves)})")
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

# ---- allocator behaviour across steps
import time
for i in range(3):
    s0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    ts.step(inp, gt_rgb, gt_mask)
    torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print(f"step {i}: {1e3 * (time.perf_counter() - t0):.1f} ms; cudaMalloc calls {s1['num_device_alloc'] - s0['num_device_alloc']}, cudaFree calls {s1['num_device_free'] - s0['num_device_free']}, "
          f"alloc retries {s1['num_alloc_retries'] - s0['num_alloc_retries']}, reserved {s1['reserved_bytes.all.current'] / 2**30:.1f} GiB, peak allocated {s1['allocated_bytes.all.peak'] / 2**30:.1f} GiB")
t0 = time.perf_counter()
xs = [torch.empty(125440, 256, device=dev) for _ in range(200)]
torch.cuda.synchronize()
print(f"200 x torch.empty(125440, 256): {1e3 * (time.perf_counter() - t0):.1f} ms")
