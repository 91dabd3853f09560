"""Round-2 starter: one process, every tcgen05 SDF-kernel variant — time per launch (sampler-round shape and the shading
shape with gradient) and max error against the exact-fp32 CUDA-core kernel.  Variants are selected per launch through the
environment (HOLD_TC_PAIR, HOLD_TC_LEAN, HOLD_TC_DBG are read at launch time).
usage: python tools/exp_matrix.py [points_log2=23]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import capi, scene_io, synth

VARIANTS = [
    ("single-CTA (default)", dict()),
    ("single-CTA LEAN", dict(HOLD_TC_LEAN="1")),
    ("single-CTA LEAN, t-stash", dict(HOLD_TC_LEAN="1", HOLD_TC_DBG="64")),
    ("single-CTA FAST", dict(HOLD_TC_FAST="1")),
    ("single-CTA FAST, 8 weight replicas", dict(HOLD_TC_FAST="1", HOLD_TC_WCOPIES="8")),
    ("pair fine hand-offs", dict(HOLD_TC_PAIR="1")),
    ("pair coarse hand-offs", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="32")),
    ("pair fine, light arrive", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="128")),
    ("pair coarse, light arrive", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="160")),
    ("pair coarse, light, 8 replicas", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="160", HOLD_TC_WCOPIES="8")),
    ("pair wide epilogue, light", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="384")),
    ("pair wide, light, 8 replicas", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="384", HOLD_TC_WCOPIES="8")),
]
ctx = capi.Context(0); dev = torch.device("cuda", 0)
sc = synth.make_scene(H=8, W=8, S=128, nodes=("right", "object"))
P = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 23)
Pg = min(P, 1 << 20)                                   # shading chunks are <= 1 Mi points
g = torch.Generator().manual_seed(0)
xc = ((torch.rand(P, 3, generator=g) - 0.5) * 1.6).to(dev)
n32 = scene_io.build_net(sc, ctx, capi.MLP_FP32).nodes["right"]
Pr = 1 << 16                                           # reference subset (the fp32 kernel is slow)
ref_sdf = torch.empty(Pr, device=dev); ref_g = torch.empty(Pr, 3, device=dev); ref_f = torch.empty(Pr, 256, device=dev)
capi.check(capi.lib().hold_sdf_eval(ctx.h, n32.slot, Pr, capi.ptr(xc), None, capi.ptr(ref_sdf), capi.ptr(ref_g), capi.ptr(ref_f), capi.stream_ptr()))
torch.cuda.synchronize()
ref = (ref_sdf.clone(), ref_g.clone(), ref_f.clone())
node = scene_io.build_net(sc, ctx, capi.MLP_TC).nodes["right"]
sdf = torch.empty(P, device=dev); grad = torch.empty(Pg, 3, device=dev); feat = torch.empty(Pg, 256, device=dev)

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1.0)).item()
print(f"{'variant':28s} {'sdf-only ms':>12s} {'TFLOP/s':>8s} {'rev ms/Mi':>10s} {'e_sdf':>9s} {'e_grad':>9s} {'e_feat':>9s}")
for name, env in VARIANTS:
    for k in ("HOLD_TC_PAIR", "HOLD_TC_LEAN", "HOLD_TC_DBG", "HOLD_TC_FAST", "HOLD_TC_WCOPIES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        t0 = timed(lambda: capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(xc), None, capi.ptr(sdf), None, None, capi.stream_ptr())))
        ctx.check()
        e0 = rel(sdf[:Pr], ref[0])
        t1 = timed(lambda: capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, Pg, capi.ptr(xc), None, capi.ptr(sdf), capi.ptr(grad), capi.ptr(feat), capi.stream_ptr())))
        ctx.check()
        print(f"{name:28s} {t0:12.2f} {2 * 459008 * P / t0 / 1e9:8.1f} {t1 * (1 << 20) / Pg:10.2f} {e0:9.2e} {rel(grad[:Pr], ref[1]):9.2e} {rel(feat[:Pr], ref[2]):9.2e}", flush=True)
    except Exception as ex:
        print(f"{name:28s} FAILED: {ex}", flush=True)


# ------------------------------------------------------------------ whole foreground step (all kernels), small frame
STEP_VARIANTS = [
    ("default", dict()),
    ("KNN filtered scan", dict(HOLD_KNN_FILTER="1")),
    ("KNN 6 blocks/SM", dict(HOLD_KNN_OCC="1")),
    ("LEAN SDF kernels", dict(HOLD_TC_LEAN="1")),
    ("LEAN + t-stash", dict(HOLD_TC_LEAN="1", HOLD_TC_DBG="64")),
    ("LEAN + KNN filter", dict(HOLD_TC_LEAN="1", HOLD_KNN_FILTER="1")),
    ("FAST + KNN filter", dict(HOLD_TC_FAST="1", HOLD_KNN_FILTER="1")),
]
sc2 = synth.make_scene(H=192, W=192, S=128, nodes=("right", "object"), B=1, seed=0)
for nid in sc2.node_ids:
    sc2.beta[nid] = torch.tensor(0.03)
net2 = scene_io.build_net(sc2, ctx, capi.MLP_TC)
inp2 = scene_io.scene_input(sc2, dev)
base = None
print(f"\n{'foreground step 192x192':28s} {'ms':>9s} {'k rays/s':>9s} {'max|d fg_rgb| vs default':>26s}")
for name, env in STEP_VARIANTS:
    for k in ("HOLD_TC_PAIR", "HOLD_TC_LEAN", "HOLD_TC_DBG", "HOLD_TC_FAST", "HOLD_TC_WCOPIES", "HOLD_KNN_FILTER", "HOLD_KNN_OCC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        out = {}
        def step():
            out["o"] = net2.forward_fg(inp2, return_factors=False)
        t = timed(step, n=2)
        ctx.check()
        rgb = out["o"]["fg_rgb"].clone()
        if base is None:
            base = rgb
        print(f"{name:28s} {t:9.1f} {192 * 192 / t:9.1f} {(rgb - base).abs().max().item():26.2e}", flush=True)
    except Exception as ex:
        print(f"{name:28s} FAILED: {ex}", flush=True)
