"""Pair kernel (HOLD_TC_PAIR=1) against the exact-fp32 CUDA-core kernel: error maps by row block / column block."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HOLD_TC_PAIR", "1")
import torch
from hold_b200 import capi, scene_io, synth

ctx = capi.Context(0); dev = torch.device("cuda", 0)
sc = synth.make_scene(H=8, W=8, S=32, nodes=("right", "object"), seed=4)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x = ((torch.rand(P, 3, generator=torch.Generator().manual_seed(9)) - 0.5) * 1.6).to(dev)
res = {}
for mode in (capi.MLP_FP32, capi.MLP_TC):
    net = scene_io.build_net(sc, ctx, mode)
    node = net.nodes["right"]
    sdf0 = torch.full((P,), float("nan"), device=dev)
    capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(x), None, capi.ptr(sdf0), None, None, capi.stream_ptr()))
    torch.cuda.synchronize()
    try:
        ctx.check()
    except Exception as ex:
        print("mode", mode, "sdf-only:", ex)
    sdf = torch.full((P,), float("nan"), device=dev); grad = torch.full((P, 3), float("nan"), device=dev); feat = torch.full((P, 256), float("nan"), device=dev)
    capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(x), None, capi.ptr(sdf), capi.ptr(grad), capi.ptr(feat), capi.stream_ptr()))
    torch.cuda.synchronize()
    try:
        ctx.check()
    except Exception as ex:
        print("mode", mode, "rev:", ex)
    res[mode] = (sdf0.cpu(), sdf.cpu(), grad.cpu(), feat.cpu())
a, b = res[capi.MLP_TC], res[capi.MLP_FP32]
for name, u, v in (("sdf_only", a[0], b[0]), ("rev.sdf", a[1], b[1]), ("rev.grad", a[2], b[2]), ("rev.feat", a[3], b[3])):
    d = (u - v).abs()
    nan = torch.isnan(u).float().mean().item()
    d = torch.nan_to_num(d, nan=9.9)
    print(f"{name}: max {d.max().item():.3e} mean {d.mean().item():.3e} nan-frac {nan:.3f}  ref-scale {v.abs().max().item():.3e}")
    n = (P // 256) * 256
    if n:
        dd = d[:n].reshape(n // 256, 8, 32, -1)   # [super-tile, 32-row block (tile X rank0 a,b | X rank1 a,b | Y ...), row, cols]
        print("   by 32-row block of a super-tile:", " ".join(f"{dd[:, k].max().item():.1e}" for k in range(8)))
    if d.dim() == 2 and d.shape[1] == 256:
        print("   by 32-column block:", " ".join(f"{d[:, 32 * k: 32 * k + 32].max().item():.1e}" for k in range(8)))
print("first rows sdf_only tc vs fp32:", a[0][:6].tolist(), b[0][:6].tolist())
