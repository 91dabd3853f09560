"""are cheaper MMA modes admissible for the sampler rounds (their SDF values feed only a PDF; the final samples
are re-evaluated at full precision by the shading pass)?  For passes = 3 (production), 2 (drop hi*lo) and 1 (fp16 x fp16 only) of
the sampler-round SDF launches: time per launch at bench size, frame time, and the pixel-level deviation of the per-node renders
from the oracle on the golden scenes next to the production path's own deviation."""
import ctypes as C
import glob
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import capi, scene_io, synth

ctx = capi.Context(0); dev = torch.device("cuda", 0)
L = capi.lib()
L.hold_debug_set.restype = C.c_int
L.hold_debug_set.argtypes = [C.c_void_p, C.c_int, C.c_int]

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- parity on the golden scenes (reference-generated fixtures)
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "*.pt")))
print(f"{'golden':22s} {'passes':>6s} {'min share of pixels within 1e-4 (per node)':>44s} {'max |d|':>10s} {'mean |d|':>10s}")
for path in GOLD:
    rec = torch.load(path)
    sc = synth.make_scene(**rec["scene_kwargs"])
    for nid in sc.node_ids:
        sc.beta[nid] = torch.tensor(rec["beta"])
    net = scene_io.build_net(sc, ctx, capi.MLP_TC)
    for passes in (3, 2, 1):
        L.hold_debug_set(ctx.h, 3, passes)
        out = net.forward_fg(scene_io.scene_input(sc, dev, ray_ids=rec["ray_ids"]))
        ctx.check()
        fr, mx, mn = 1.0, 0.0, 0.0
        for nid in sc.node_ids:
            for k in ("fg_rgb", "mask_prob", "depth", "normal"):
                b = rec["render"][nid][k].float()
                a = out[f"{nid}.{k}"].detach().float().cpu().reshape(b.shape)
                d = (a - b).abs()
                sca = max(1.0, b.abs().max().item())
                fr = min(fr, (d <= 1e-4 * sca).float().mean().item()); mx = max(mx, d.max().item() / sca); mn = max(mn, d.mean().item() / sca)
        print(f"{os.path.basename(path)[:-3]:22s} {passes:6d} {fr:44.4f} {mx:10.2e} {mn:10.2e}", flush=True)
L.hold_debug_set(ctx.h, 3, 3)

# ---- time: one sampler-round SDF launch inside hold_sample, and the frame
import bench
sc = bench.make_scene(0)
net = scene_io.build_net(sc, ctx, capi.MLP_TC)
inp = scene_io.scene_input(sc, dev)
for passes in (3, 2, 1):
    L.hold_debug_set(ctx.h, 3, passes)
    t = timed(lambda: net.forward_fg(inp, return_factors=False, want_weights=False), n=2)
    ctx.check()
    print(f"frame 512x512, sampler rounds with {passes} pass(es): {t:8.1f} ms  ({512 * 512 / t:6.1f} k rays/s)", flush=True)
L.hold_debug_set(ctx.h, 3, 3)
print("exp_sampler_precision done")
