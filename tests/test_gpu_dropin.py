"""The reference-signature adaptors (hold_b200/dropin.py) called the way the reference's call sites call them
(mano_node.py:100-109, node.py:57-67, volsdf_utils.py:143-146,150-169), against the oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1.0)).item()


@pytest.fixture(scope="module")
def env(ctx):
    from hold_b200 import capi, scene_io, synth
    from oracle import hold_oracle as O

    sc = synth.make_scene(H=12, W=12, S=32, nodes=("right", "object"), B=2, seed=5)
    sc.sampler["add_tiny"] = 1e-3   # whole-loop z_vals are compared: see tests/test_gpu_stages.py on the default constant
    net = scene_io.build_net(sc, ctx, capi.MLP_FP32)
    dev = torch.device("cuda", 0)
    return dict(sc=sc, net=net, dev=dev, art=O.scene_articulation(sc), O=O, inp=scene_io.scene_input(sc, dev))


def test_sampler_with_the_reference_call(env, ctx):
    from hold_b200 import dropin
    from hold_b200.model import ErrorBoundSampler as Mirror

    sc, O, dev = env["sc"], env["O"], env["dev"]
    dirs, cam = O.camera_rays(sc.uv, sc.extrinsics, sc.intrinsics)
    P = dirs.shape[1]
    dirs = dirs.reshape(-1, 3).to(dev)
    cam = cam.unsqueeze(1).repeat(1, P, 1).reshape(-1, 3).to(dev)
    for nid in sc.node_ids:
        node, a = env["net"].nodes[nid], env["art"][nid]
        smp = dropin.ErrorBoundSampler(sc.bounding_sphere, inverse_sphere_bg=True, **{k: v for k, v in sc.sampler.items()})
        deformer = dropin.MANODeformer(node) if nid != "object" else dropin.ObjectDeformer(node)
        # the servers' own outputs (what mano_node.py:72-98 / object_node.py:58-82 put into deform_info)
        pose, keep, srv, tfs = node.articulate(env["inp"])
        if nid != "object":
            deform_info = {"cond": {"pose": a["pose_cond"].to(dev)}, "tfs": srv["tfs"], "verts": srv["verts"]}
        else:
            deform_info = {"cond": {"pose": torch.zeros(sc.B, 0, device=dev)}, "tfs": srv["obj_tfs"]}
        z = smp.get_z_vals(dropin.sdf_func_with_deformer, deformer, node.implicit_network, dirs, cam, node.density, False, deform_info)
        ctx.check()
        z2, it2 = Mirror(node).get_z_vals(dirs, cam, pose, sc.B)
        assert torch.equal(z, z2), f"{nid}: adaptor and mirror run the same kernels on the same inputs"
        assert int(smp.last_iters.item()) == int(it2.item())
        assert smp.inverse_sphere_sampler.inverse_sample(dirs, cam, False, sc.bounding_sphere).shape == (dirs.shape[0], 32)


def test_deformers_and_sdf_func(env, ctx):
    from hold_b200 import dropin

    sc, O, dev = env["sc"], env["O"], env["dev"]
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(sc.B, 700, 3, generator=g) - 0.5) * 2.0
    for nid in sc.node_ids:
        node, a = env["net"].nodes[nid], env["art"][nid]
        hand = nid != "object"
        d = dropin.MANODeformer(node) if hand else dropin.ObjectDeformer(node)
        tfs = a["tfs"].to(dev)
        verts = a["verts"].to(dev) if hand else None
        xc, mask = d.forward(x.to(dev), tfs, return_weights=False, inverse=True, verts=verts)
        xd = d.forward_skinning(x.to(dev) * 0.3, None, tfs)
        ctx.check()
        for b in range(sc.B):
            if hand:
                xo, mo, _ = O.hand_inverse_warp(x[b], a["verts"][b], a["skin_W"], a["tfs"][b])
                assert rel(xc[b], xo) < 1e-4 and (mask[b].cpu() == mo).float().mean().item() > 0.999
                w, _, _ = O.skin_weights_query(x[b] * 0.3, a["cano_verts"], a["skin_W"])
                T = torch.einsum("pn,nij->pij", w, a["tfs"][b])
                xdo = torch.einsum("pij,pj->pi", T, F.pad(x[b] * 0.3, (0, 1), value=1.0))[:, :3]
            else:
                assert mask is None
                assert rel(xc[b], O.rigid_inverse_warp(x[b], a["tfs"][b])) < 1e-4
                xdo = (F.pad(x[b] * 0.3, (0, 1), value=1.0) @ a["tfs"][b].T)[:, :3]
            assert rel(xd[b], xdo) < 1e-4, f"{nid}: forward_skinning {rel(xd[b], xdo):.2e}"
        # sdf_func_with_deformer(deformer, sdf_fn, training, x, deform_info) -> (sdf, x_c, feature)
        info = {"cond": {"pose": None}, "tfs": tfs}
        if hand:
            info["verts"] = verts
        sdf, x_c, feat = dropin.sdf_func_with_deformer(d, node.implicit_network, False, x.to(dev).reshape(-1, 3), info)
        assert sdf.shape == (sc.B, 700, 1) and feat.shape == (sc.B, 700, 256)
        ref = O.sdf_mlp(x_c[0].cpu(), sc.sdf_state[nid], torch.zeros(700, 45) if hand else None)
        assert rel(sdf[0, :, 0], ref[:, 0]) < 1e-4 and rel(feat[0], ref[:, 1:]) < 1e-4


def test_rendering_network_call(env, ctx):
    from hold_b200 import dropin

    sc, O, dev = env["sc"], env["O"], env["dev"]
    g = torch.Generator().manual_seed(4)
    n = 500
    for nid in sc.node_ids:
        node, a = env["net"].nodes[nid], env["art"][nid]
        hand = nid != "object"
        pts = (torch.rand(sc.B * n, 3, generator=g) - 0.5)
        nrm = F.normalize(torch.randn(sc.B * n, 3, generator=g), dim=1)
        feat = torch.randn(sc.B * n, 256, generator=g) * 0.3
        fr = torch.arange(sc.B).repeat_interleave(n)
        if hand:
            pose = a["pose_cond"]
            out = dropin.RenderingNetAdaptor(node)(pts.to(dev), nrm.to(dev), None, pose.to(dev), feat.to(dev))
            ref = O.rgb_mlp(pts, nrm, pose[fr], feat, sc.rgb_state[nid])
        else:
            tc = sc.time_code
            fv = torch.cat([feat, tc[fr]], -1)
            out = dropin.RenderingNetAdaptor(node)(pts.to(dev), nrm.to(dev), None, torch.zeros(sc.B, 0, device=dev), fv.to(dev))
            ref = O.rgb_mlp(pts, nrm, None, fv, sc.rgb_state[nid])
        ctx.check()
        assert rel(out, ref) < 1e-4, f"{nid}: {rel(out, ref):.2e}"


def test_constructors_from_reference_config(ctx):
    """dropin.HOLDNet(opt, betas_r, betas_l, num_frames, args) built from the `model:` block of confs/general.yaml renders the same
    pixels as the mirror assembled by hand (same weights), and the per-node objects are hold_b200.model.Node instances."""
    import copy

    from hold_b200 import capi, dropin, scene_io, synth

    sc = synth.make_scene(H=8, W=8, S=32, nodes=("right", "object"), B=2, seed=4)
    sc.intrinsics[:, 0, 2] += 0.37
    dev = torch.device("cuda", 0)
    opt = copy.deepcopy(dropin._SUPPORTED)
    opt["rendering_network"]["d_in"] = 14
    opt["density"] = {"params_init": {"beta": 0.1}, "beta_min": 1e-4}
    opt["ray_sampler"] = dict(sc.sampler, N_samples_inverse_sphere=32)
    opt["scene_bounding_sphere"] = sc.bounding_sphere
    args = {"n_images": sc.B, "case": "synthetic"}
    net = dropin.HOLDNet(opt, sc.betas["right"], None, sc.B, args, ctx=ctx, mano_r=sc.mano["right"], obj_pts=sc.obj_pts_cano, mlp_mode=capi.MLP_TC).to(dev)
    assert list(net.nodes) == ["right", "object"] and net.background is not None
    assert opt["rendering_network"]["d_in"] == 14 + 32     # the reference's in-place edit (object_node.py:19)
    ref = scene_io.build_net(sc, ctx, capi.MLP_TC)
    ref_sd = {nid: {k: v.clone() for k, v in ref.nodes[nid].state_dict().items()} for nid in ref.nodes}
    for nid, node in net.nodes.items():
        node.load_state_dict(ref_sd[nid], strict=True)
        node.sync_weights()
    bg, _, _ = scene_io.build_background(sc, ctx, mlp_mode=capi.MLP_TC)
    bg_sd = {k: v.clone() for k, v in bg.state_dict().items()}
    inp = scene_io.scene_input(sc, dev)
    net.background.load_state_dict(bg_sd, strict=True)
    net.background.sync_weights()
    a = net(inp)
    from hold_b200.model import HOLDNet
    for node in ref.nodes.values():   # the two nets share ctx slots: re-upload the hand-assembled one's weights before using it
        node.sync_weights()
    bg.sync_weights()
    b = HOLDNet(ctx, dict(ref.nodes), background=bg)(inp)
    ctx.check()
    for k in ("rgb", "fg_rgb", "depth", "normal", "semantics"):
        assert torch.equal(a[k], b[k]), k
