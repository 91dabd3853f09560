"""the NeRF++ background leg and the final composite against the oracle (which matches the
reference's Background class bit for bit, oracle/ref_harness.py check_background)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_background_and_full_composite(ctx):
    from hold_b200 import capi, scene_io, synth
    from hold_b200.model import HOLDNet
    from oracle import hold_oracle as O

    sc = synth.make_scene(H=10, W=10, S=32, nodes=("right", "object"), B=2, seed=8)
    sc.intrinsics[:, 0, 2] += 0.37   # keep every ray off the sphere centre (0/0 in the reference's depth2pts_outside)
    sc.intrinsics[:, 1, 2] -= 0.21
    dev = torch.device("cuda", 0)
    net = scene_io.build_net(sc, ctx, capi.MLP_FP32)
    bg, sdf_sd, rgb_sd = scene_io.build_background(sc, ctx)
    full = HOLDNet(ctx, dict(net.nodes), background=bg)
    out = full(scene_io.scene_input(sc, dev))
    ctx.check()
    # oracle: foreground, then the background leg on the oracle's own bg_weights
    ref, _ = O.render_scene(sc, stable_ties=True)
    dirs, cam = O.camera_rays(sc.uv, sc.extrinsics, sc.intrinsics)
    P = dirs.shape[1]
    dirs, cam = dirs.reshape(-1, 3), cam.unsqueeze(1).repeat(1, P, 1).reshape(-1, 3)
    frame = torch.arange(sc.B).repeat_interleave(P)
    fc = bg.frame_latent_encoder.weight.data.cpu()[sc.frame_idx]
    # stage parity: same bg_weights in
    bgw = out["bg_weights"].cpu()
    o = O.background(bgw, dirs, cam, fc, frame, sdf_sd, rgb_sd, sc.bounding_sphere)
    for name, a, b in (("bg_rgb", out["rgb"].cpu() - out["fg_rgb"].cpu(), o[0]), ("bg_rgb_only", out["bg_rgb_only"].cpu(), o[1]),
                       ("bg_z_vals", out["bg_z_vals"].cpu(), o[3])):
        err = (a - b).abs().max().item()
        print(f"{name}: max|d| {err:.2e}")
        assert err <= 1e-5, f"{name}: max|d| {err:.2e}"
    sem = out["semantics"].cpu() - out["fg_semantics"].cpu()
    assert (sem - o[2]).abs().max().item() <= 1e-6
    assert out["instance_map"].shape == (sc.B * P,) and out["instance_map"].dtype == torch.int64
    # end to end: final rgb against oracle fg + oracle bg on the oracle's bg_weights (composite noise floor)
    assert len(ref) == 1
    r = ref[0]["render"]["comp"]
    ob = O.background(r["bg_weights"], dirs, cam, fc, frame, sdf_sd, rgb_sd, sc.bounding_sphere)
    d = (out["rgb"].cpu() - (r["fg_rgb"] + ob[0])).abs()
    assert d.mean().item() <= 8e-3 and d.max().item() <= 1.5e-1
