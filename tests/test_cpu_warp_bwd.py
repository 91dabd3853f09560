"""CPU-only: the inverse-warp backward the CUDA kernels run (hold_b200/csrc/warp_bwd_phases.h, compiled for the host) against
torch.autograd over the oracle's inverse warps — gradients w.r.t. the bone / object transforms and the posed points."""
import ctypes as C
import os
import subprocess

import torch

from hold_b200 import synth
from oracle import hold_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    out = os.path.join(ROOT, "tests", "_build", "libwarp_bwd_host.so")
    src = os.path.join(ROOT, "tests", "host", "warp_bwd_host.cpp")
    hdr = os.path.join(ROOT, "hold_b200", "csrc", "warp_bwd_phases.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


_p = lambda t: None if t is None else C.c_void_p(t.data_ptr())


def test_hand_inverse_warp_backward_on_host():
    lib = _lib()
    sc = synth.make_scene(H=4, W=4, S=32, nodes=("right",), B=1, seed=3)
    art = O.scene_articulation(sc)["right"]
    verts, tfs, W = art["verts"][0].contiguous(), art["tfs"][0].contiguous(), art["skin_W"].contiguous()
    g = torch.Generator().manual_seed(0)
    P = 300
    x = (verts[torch.randint(0, 778, (P,), generator=g)] + 0.05 * torch.randn(P, 3, generator=g)).contiguous()
    x[:20] = x[:20] + 1.5                      # far points: clamped confidences
    gxc = torch.randn(P, 3, generator=g)
    xr, tr = x.clone().requires_grad_(), tfs.clone().requires_grad_()
    xc, _, idx = O.hand_inverse_warp(xr, verts, W, tr)
    # the reference detaches the skinning weights (deformer.py:101): x only enters through the homogeneous point
    w, _, _ = O.skin_weights_query(x, verts, W)
    T = torch.einsum("pn,nij->pij", w.detach(), tr)
    xc2 = torch.einsum("pij,pj->pi", T.inverse(), torch.nn.functional.pad(xr, (0, 1), value=1.0))[:, :3]
    ref_t, ref_x = torch.autograd.grad((xc2 * gxc).sum(), (tr, xr))
    for nt in (32, 128):
        g_x = torch.full((P, 3), float("nan"))
        g_t = torch.full((16, 4, 4), float("nan"))
        assert lib.warp_bwd_hand_host(C.c_int(nt), C.c_int(P), _p(x), _p(idx.to(torch.int32).contiguous()), _p(verts), _p(W), _p(tfs),
                                      _p(gxc), _p(g_x), _p(g_t)) == 0
        # autograd through the full 4x4 inverse also assigns gradient to tfs[:, 3, :3]; the servers keep that row constant
        # (0, 0, 0, 1) and hold_mano_lbs_bwd ignores it: compare the entries that carry meaning
        mask = torch.ones(16, 4, 4, dtype=torch.bool)
        mask[:, 3, :3] = False
        for name, a, b in (("g_tfs", g_t[mask], ref_t[mask]), ("g_x", g_x, ref_x)):
            err = (a - b).abs().max().item() / max(1.0, b.abs().max().item())
            assert err < 2e-5, f"nt={nt} {name}: {err:.2e}"


def test_object_inverse_warp_backward_on_host():
    lib = _lib()
    g = torch.Generator().manual_seed(1)
    P = 257
    tf = torch.eye(4)
    tf[:3, :3] = O.axis_angle_to_matrix(torch.tensor([[0.3, -0.5, 0.2]]))[0] * 1.3
    tf[:3, 3] = torch.tensor([0.1, -0.2, 0.3])
    x, gxc = torch.randn(P, 3, generator=g), torch.randn(P, 3, generator=g)
    xr, tr = x.clone().requires_grad_(), tf.clone().requires_grad_()
    ref_t, ref_x = torch.autograd.grad((O.rigid_inverse_warp(xr, tr) * gxc).sum(), (tr, xr))
    g_x, g_t = torch.full((P, 3), float("nan")), torch.full((4, 4), float("nan"))
    assert lib.warp_bwd_obj_host(C.c_int(64), C.c_int(P), _p(x.contiguous()), _p(tf.contiguous()), _p(gxc), _p(g_x), _p(g_t)) == 0
    # the reference inverts the full 4x4: entries of the last row other than [3][3] also receive autograd gradient there, but
    # the servers keep that row constant (0, 0, 0, s): compare the entries that carry meaning
    mask = torch.ones(4, 4, dtype=torch.bool)
    mask[3, :3] = False
    assert ((g_t - ref_t)[mask].abs().max() / ref_t.abs().max().clamp_min(1.0)).item() < 2e-5
    assert ((g_x - ref_x).abs().max() / ref_x.abs().max().clamp_min(1.0)).item() < 2e-5
