"""Cycle accounting of the pair kernel (HOLD_TC_PAIR=1 HOLD_TC_PROF=1) under the HOLD_TC_DBG experiment switches."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import capi, scene_io, synth
ctx = capi.Context(0); dev = torch.device("cuda", 0)
sc = synth.make_scene(H=8, W=8, S=128, nodes=("right", "object"))
net = scene_io.build_net(sc, ctx, capi.MLP_TC)
P = 65536 * 128
xc = (torch.rand(P, 3, device=dev) - 0.5) * 1.6
sdf = torch.empty(P, device=dev)
node = net.nodes["right"]
L = capi.lib()
L.hold_debug_ws_copy.restype = C.c_int
L.hold_debug_ws_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
def launch():
    capi.check(L.hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(xc), None, capi.ptr(sdf), None, None, capi.stream_ptr()))
for tok in (sys.argv[1:] or ["1:0"]):
    pair, dbg = (int(v) for v in tok.split(":"))
    os.environ["HOLD_TC_PAIR"] = str(pair)
    os.environ["HOLD_TC_DBG"] = str(dbg)
    launch(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): launch()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    try:
        ctx.check()
    except Exception as ex:
        print("check:", ex)
    line = f"pair={int(pair)} dbg={dbg:2d}: {ms:7.2f} ms  ({ms * 4:6.1f} ms at bench size)"
    if not pair and os.environ.get("HOLD_TC_PROF"):
        t = torch.zeros(64, dtype=torch.int64, device=dev)
        assert L.hold_debug_ws_copy(ctx.h, 23, C.c_void_p(t.data_ptr()), C.c_size_t(64 * 8)) == 0
        v = t.cpu().tolist()
        tot = max(v[0], 1)
        line += (f"\n    single-CTA kernel, CTA 0: mma warp waits hand-offs {v[1]/tot:.2f} weights {v[2]/tot:.2f};"
                 f" producer waits free stage {v[9]/max(v[8],1):.2f}; epilogue waits accumulator w2 {v[17]/max(v[16],1):.2f} w17 {v[21]/max(v[20],1):.2f}")
    if pair and os.environ.get("HOLD_TC_PROF"):
        t = torch.zeros(64, dtype=torch.int64, device=dev)
        assert L.hold_debug_ws_copy(ctx.h, 23, C.c_void_p(t.data_ptr()), C.c_size_t(64 * 8)) == 0
        v = t.cpu().tolist()
        tot = max(v[0], 1)
        line += (f"\n    mma: total {v[0]/1e6:.2f} Mclk  wait a_ready {v[1]/tot:.2f} w_full {v[2]/tot:.2f}"
                 f"\n    producer leader: wait w_empty {v[9]/max(v[8],1):.2f}   peer: {v[11]/max(v[10],1):.2f}   forwarder wait w_full {v[12]/tot:.2f}"
                 f"\n    epilogue d_full wait: leader w2 {v[17]/max(v[16],1):.2f} w17 {v[21]/max(v[20],1):.2f}  peer w2 {v[25]/max(v[24],1):.2f} w17 {v[29]/max(v[28],1):.2f}")
    print(line, flush=True)
