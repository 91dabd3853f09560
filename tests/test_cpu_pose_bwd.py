"""CPU-only: the closed-form reverse mode the pose-server backward kernels implement (oracle/manual_bwd.py) against
torch.autograd over the oracle's forward — which is how the reference itself differentiates these servers
(fitting/model.py:117).  float64 so that the comparison checks the derivation, not rounding."""
import torch

from hold_b200 import synth
from oracle import hold_oracle as O
from oracle import manual_bwd as MB


def _mano64(seed=0):
    m = synth.make_mano_struct(seed)
    return {k: (v.double() if torch.is_tensor(v) and v.dtype.is_floating_point else v) for k, v in m.items()}


def test_mano_server_backward_matches_autograd():
    m = _mano64()
    g = torch.Generator().manual_seed(3)
    B = 2
    betas = (0.5 * torch.randn(B, 10, generator=g)).double().requires_grad_()
    pose = (0.4 * torch.randn(B, 48, generator=g)).double().requires_grad_()
    transl = torch.randn(B, 3, generator=g).double().requires_grad_()
    scale = (1.0 + 0.3 * torch.rand(B, generator=g)).double().requires_grad_()
    _, tci = O.mano_canonical({k: (v.float() if torch.is_tensor(v) and v.dtype == torch.float64 else v) for k, v in m.items()}, betas[0].detach().float())
    tci = tci.double()
    gv = torch.randn(B, 778, 3, generator=g).double()
    gj = torch.randn(B, 21, 3, generator=g).double()
    gt = torch.randn(B, 16, 4, 4, generator=g).double()
    for tc in (None, tci):
        out = O.mano_server(m, scale, transl, pose, betas, tc)
        loss = (out["verts"] * gv).sum() + (out["jnts"] * gj).sum() + (out["tfs"] * gt).sum()
        ref = torch.autograd.grad(loss, (betas, pose, transl, scale))
        got = MB.mano_server_bwd(m, scale.detach(), transl.detach(), pose.detach(), betas.detach(), tc, gv, gj, gt)
        for name, a, b in zip(("betas", "pose", "transl", "scale"), got, ref):
            err = (a - b).abs().max().item() / max(1.0, b.abs().max().item())
            assert err < 1e-9, f"{name} (tfs_c_inv={'set' if tc is not None else 'None'}): {err:.2e}"
    # verts-only upstream gradient (what the silhouette loss of optimize_ckpt.py produces)
    out = O.mano_server(m, scale, transl, pose, betas, tci)
    ref = torch.autograd.grad((out["verts"] * gv).sum(), (betas, pose, transl, scale))
    got = MB.mano_server_bwd(m, scale.detach(), transl.detach(), pose.detach(), betas.detach(), tci, gv, None, None)
    for a, b in zip(got, ref):
        assert (a - b).abs().max().item() / max(1.0, b.abs().max().item()) < 1e-9


def test_object_server_backward_matches_autograd():
    g = torch.Generator().manual_seed(5)
    B, Nv = 3, 200
    rot = (0.8 * torch.randn(B, 3, generator=g)).double()
    rot[2] = 0.0                                    # the small-angle branch (|a| < 1e-6)
    rot.requires_grad_()
    trans = torch.randn(B, 3, generator=g).double().requires_grad_()
    ss = (1.0 + 0.3 * torch.rand(B, generator=g)).double().requires_grad_()
    osc = torch.tensor(1.7, dtype=torch.float64, requires_grad=True)
    D = torch.eye(4, dtype=torch.float64)
    D[:3, :3] *= 0.8
    D[:3, 3] = torch.tensor([0.1, -0.2, 0.05], dtype=torch.float64)
    pts = torch.randn(Nv, 3, generator=g).double()
    gv = torch.randn(B, Nv, 3, generator=g).double()
    gt = torch.randn(B, 4, 4, generator=g).double()
    tf, v = O.object_server(rot, trans, ss, osc, D, pts)
    ref = torch.autograd.grad((v * gv).sum() + (tf * gt).sum(), (rot, trans, ss, osc))
    got = MB.object_server_bwd(rot.detach(), trans.detach(), ss.detach(), osc.detach(), D, pts, gv, gt)
    for name, a, b in zip(("rot", "trans", "scene_scale", "obj_scale"), got, ref):
        a, b = torch.nan_to_num(a), torch.nan_to_num(b)   # autograd of norm() at exactly 0 is 0/0-free in torch; keep symmetric
        err = (a - b).abs().max().item() / max(1.0, b.abs().max().item())
        assert err < 1e-9, f"{name}: {err:.2e}"


# ------------------------------------------------------------------ the kernels' own phase functions, run on the host
def _host_lib():
    import ctypes as C
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "_build", "libpose_bwd_host.so")
    src = os.path.join(root, "tests", "host", "pose_bwd_host.cpp")
    hdr = os.path.join(root, "hold_b200", "csrc", "pose_bwd_phases.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def _p(t):
    import ctypes as C

    return None if t is None else C.c_void_p(t.data_ptr())


def test_mano_backward_kernel_phases_on_host():
    """hold_b200/csrc/pose_bwd_phases.h — the code k_mano_lbs_bwd runs — executed tid by tid on the CPU (fp32) against
    torch.autograd over the oracle (fp32)."""
    import ctypes as C

    lib = _host_lib()
    m = synth.make_mano_struct(1)
    g = torch.Generator().manual_seed(13)
    B = 3
    betas = 0.5 * torch.randn(B, 10, generator=g)
    pose = 0.4 * torch.randn(B, 48, generator=g)
    transl = torch.randn(B, 3, generator=g)
    scale = 1.0 + 0.3 * torch.rand(B, generator=g)
    _, tci = O.mano_canonical(m, betas[0])
    gv, gj, gt = torch.randn(B, 778, 3, generator=g), torch.randn(B, 21, 3, generator=g), torch.randn(B, 16, 4, 4, generator=g)
    par = m["parents"].to(torch.int32).contiguous()
    tips = m["tip_ids"].to(torch.int32).contiguous()
    mt = {k: m[k].float().contiguous() for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "hands_mean")}
    for nt in (64, 256):
        for use in ((True, True, True, True), (True, False, False, True), (False, True, False, False), (False, False, True, True)):
            uv, uj, ut, utc = use
            leaves = [t.clone().requires_grad_() for t in (betas, pose, transl, scale)]
            out = O.mano_server(m, leaves[3], leaves[2], leaves[1], leaves[0], tci if utc else None)
            loss = (out["verts"] * gv).sum() * uv + (out["jnts"] * gj).sum() * uj + (out["tfs"] * gt).sum() * ut
            ref = torch.autograd.grad(loss, leaves)
            got = [torch.full((B, 10), float("nan")), torch.full((B, 48), float("nan")), torch.full((B, 3), float("nan")), torch.full((B,), float("nan"))]
            tc = tci.contiguous() if utc else None
            rc = lib.pose_bwd_mano_host(C.c_int(nt), C.c_int(B), _p(mt["v_template"]), _p(mt["shapedirs"]), _p(mt["posedirs"]),
                                        _p(mt["J_regressor"]), _p(mt["lbs_weights"]), _p(mt["hands_mean"]), _p(par), _p(tips),
                                        _p(betas), _p(pose), _p(transl), _p(scale), _p(tc), _p(gv if uv else None),
                                        _p(gj if uj else None), _p(gt if ut else None), _p(got[0]), _p(got[1]), _p(got[2]), _p(got[3]))
            assert rc == 0
            for name, a, b in zip(("betas", "pose", "transl", "scale"), got, ref):
                err = (a - b).abs().max().item() / max(1.0, b.abs().max().item())
                assert err < 2e-5, f"nt={nt} use={use} {name}: {err:.2e}"


def test_object_backward_kernel_phases_on_host():
    import ctypes as C

    lib = _host_lib()
    g = torch.Generator().manual_seed(17)
    B, Nv = 3, 333
    rot = 0.8 * torch.randn(B, 3, generator=g)
    rot[1] = 0.0
    trans = torch.randn(B, 3, generator=g)
    ss = 1.0 + 0.3 * torch.rand(B, generator=g)
    D = torch.eye(4)
    D[:3, :3] *= 0.8
    D[:3, 3] = torch.tensor([0.1, -0.2, 0.05])
    pts = torch.randn(Nv, 3, generator=g).contiguous()
    gv, gt = torch.randn(B, Nv, 3, generator=g), torch.randn(B, 4, 4, generator=g)
    for nt in (32, 256):
        for uv, ut in ((True, True), (True, False), (False, True)):
            leaves = [t.clone().requires_grad_() for t in (rot, trans, ss)]
            osc = torch.tensor(1.7, requires_grad=True)
            tf, v = O.object_server(leaves[0], leaves[1], leaves[2], osc, D, pts)
            ref = torch.autograd.grad((v * gv).sum() * uv + (tf * gt).sum() * ut, leaves + [osc])
            got = [torch.full((B, 3), float("nan")), torch.full((B, 3), float("nan")), torch.full((B,), float("nan")), torch.full((B,), float("nan"))]
            rc = lib.pose_bwd_object_host(C.c_int(nt), C.c_int(B), _p(rot), _p(trans), _p(ss), C.c_float(1.7), _p(D.contiguous()), _p(pts),
                                          C.c_int(Nv), _p(gv if uv else None), _p(gt if ut else None), _p(got[0]), _p(got[1]), _p(got[2]), _p(got[3]))
            assert rc == 0
            got[3] = got[3].sum()
            for name, a, b in zip(("rot", "trans", "scene_scale", "obj_scale"), got, ref):
                err = (torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max().item() / max(1.0, b.abs().max().item())
                assert err < 2e-5, f"nt={nt} {name}: {err:.2e}"
