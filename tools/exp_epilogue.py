"""Timing experiments on the tcgen05 epilogue (never part of the product build): variants of libhold_b200.so compiled with
-DHOLD_TC_EXP=n knock one ingredient out of the forward epilogue (results are then WRONG; only the time is read):
  1 no MUFU, 2 one MUFU, 3 no fence.proxy.async, 4 no hi/lo split, 5 no bias loads, 6 = 1 + 3 + 4 + 5 (accumulator -> FFMA -> store -> arrive only); 11 two weight stages instead of three, 12 hand-off loops unrolled by four; 0 = the product kernel.
  python tools/exp_epilogue.py --build      (authoring container: nvcc, writes build_exp/libhold_exp<n>.so)
  python tools/exp_epilogue.py --run n      (GPU box: times the sdf-only launch on 2^22 points)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = [0, 1, 2, 3, 4, 5, 6, 7, 11, 12]
OUT = os.path.join(ROOT, "build_exp")


def build(which=None):
    import __graft_entry__ as g
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for n in (which or VARIANTS):
        cmd = ["nvcc"] + g.NVCC_FLAGS + ([f"-DHOLD_TC_EXP={n}"] if n else []) + ["-o", os.path.join(OUT, f"libhold_exp{n}.so"), os.path.join(g.CSRC, "api.cu")]
        procs.append(subprocess.Popen(cmd, cwd=g.CSRC))
    for p in procs:
        assert p.wait() == 0


def run(n, passes=3):
    import torch
    from hold_b200 import capi
    capi.LIB_PATH = os.path.join(OUT, f"libhold_exp{n}.so")
    from hold_b200 import scene_io, synth
    ctx = capi.Context(0); dev = torch.device("cuda", 0)
    sc = synth.make_scene(H=8, W=8, S=128, nodes=("right", "object"))
    node = scene_io.build_net(sc, ctx, capi.MLP_TC).nodes["right"]
    P = 1 << 22
    xc = ((torch.rand(P, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 1.6).to(dev)
    sdf = torch.empty(P, device=dev)
    L = capi.lib()
    call = lambda: capi.check(L.hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(xc), None, capi.ptr(sdf), None, None, capi.stream_ptr()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if n == 7:   # cycle accounting of one tile of CTA 0: epilogue warp 2 (quarter 2 / sub 0 lane 0) and the MMA issuer
        buf = torch.empty(1024, dtype=torch.int32, device=dev)
        import ctypes as C
        L.hold_debug_ws_copy.restype, L.hold_debug_ws_copy.argtypes = C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        capi.check(L.hold_debug_ws_copy(ctx.h, 12, C.c_void_p(buf.data_ptr()), 4096))
        torch.cuda.synchronize()
        b = buf.cpu().numpy().astype("int64") & 0xFFFFFFFF
        t0 = b[0]
        print(f'  tile start (before the prologue): {int((b[12] - t0) & 0xFFFFFFFF) - (1 << 32)}')
        for l in range(8):
            e = (b[l * 16: l * 16 + 10] - t0) & 0xFFFFFFFF
            m = (b[256 + l * 16: 256 + l * 16 + 8] - t0) & 0xFFFFFFFF
            print(f"  layer {l}: epi wait_start {e[0]:6d} d_full {e[1]:6d} arrive " + " ".join(f"{x:6d}" for x in e[2:10]) + "   | mma stage issue " + " ".join(f"{x:6d}" for x in m))
    P1 = 1 << 20
    grad, feat = torch.empty(P1, 3, device=dev), torch.empty(P1, 256, device=dev)
    call3 = lambda: capi.check(L.hold_sdf_eval(ctx.h, node.slot, P1, capi.ptr(xc), None, capi.ptr(sdf), capi.ptr(grad), capi.ptr(feat), capi.stream_ptr()))
    for _ in range(3):
        call3()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        call3()
    e1.record(); torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1) / 10
    print(f"exp {n}: {ms:.3f} ms per 2^22 points  ({P * 0.918e6 / ms / 1e9:.1f} TFLOP/s algorithmic)   reverse mode: {ms3:.3f} ms per 2^20 points", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--build":
        build([int(x) for x in sys.argv[2:]])
    else:
        run(int(sys.argv[2]))
