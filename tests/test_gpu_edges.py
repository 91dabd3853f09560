"""Edge cases of the boundary: empty input, ragged ray counts, a ray that misses the bounding sphere (the reference
calls exit(), engine/ray_sampler.py:15-18), and training-mode sampling with the random draws passed as INPUTS."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fp32", "tc"])
def env(ctx, request):
    """Both arithmetic modes of the MLPs (the tensor-core mode is the one bench.py and smoke() run)."""
    from hold_b200 import capi, scene_io, synth
    from oracle import hold_oracle as O

    sc = synth.make_scene(H=8, W=8, S=32, nodes=("right", "object"), seed=6)
    sc.sampler["add_tiny"] = 1e-3   # see tests/test_gpu_stages.py: keeps the sampler comparison above libm noise
    net = scene_io.build_net(sc, ctx, capi.MLP_TC if request.param == "tc" else capi.MLP_FP32)
    dev = torch.device("cuda", 0)
    return dict(sc=sc, net=net, dev=dev, O=O, inp=scene_io.scene_input(sc, dev), art=O.scene_articulation(sc))


def test_empty_input_is_a_noop(env, ctx):
    from hold_b200 import capi

    node = env["net"].nodes["right"]
    pose, keep, _, _ = node.articulate(env["inp"])
    z = torch.empty(0, node.S, device=env["dev"])
    rc = capi.lib().hold_sample(ctx.h, node.slot, 0, 1, None, None, C.byref(pose), None, capi.ptr(z), None, capi.stream_ptr())
    assert rc == 0
    assert capi.lib().hold_sdf_eval(ctx.h, node.slot, 0, None, None, None, None, None, capi.stream_ptr()) == 0
    ctx.check()


def test_ray_missing_the_sphere_is_an_error_code(env, ctx):
    from hold_b200 import capi
    from hold_b200.model import ErrorBoundSampler

    node = env["net"].nodes["object"]
    pose, keep, _, _ = node.articulate(env["inp"])
    cam = torch.tensor([[20.0, 0.0, 0.0]] * 4, device=env["dev"])       # outside the radius-6 sphere ...
    dirs = torch.tensor([[0.0, 1.0, 0.0]] * 4, device=env["dev"])       # ... and the ray's LINE misses it (under_sqrt <= 0)
    ErrorBoundSampler(node).get_z_vals(dirs, cam, pose, 1)
    with pytest.raises(capi.HoldError) as e:
        ctx.check()
    assert "error -3" in str(e.value)
    ctx.check()  # the error word is cleared after being reported


@pytest.mark.parametrize("n_rays", [1, 37])
def test_ragged_ray_counts(env, ctx, n_rays):
    from hold_b200 import scene_io

    sc, O = env["sc"], env["O"]
    ids = torch.arange(n_rays) * (sc.H * sc.W // n_rays)
    out = env["net"].forward_fg(scene_io.scene_input(sc, env["dev"], ray_ids=ids))
    ctx.check()
    ref, _ = O.render_scene(sc, ray_ids=ids, stable_ties=True)
    for k, nid in enumerate(sc.node_ids):
        for key in ("fg_rgb", "depth", "mask_prob"):
            a, b = out[f"{nid}.{key}"].cpu().reshape(ref[0]["render"][k][key].shape), ref[0]["render"][k][key]
            assert (a - b).abs().max().item() < 1e-3, f"{nid}.{key} with {n_rays} rays"


def test_training_mode_sampling_with_given_randomness(env, ctx):
    """Stratified jitter (ray_sampler.py:70-78), random u (:292) and the extras permutation (:328) are inputs."""
    from hold_b200.model import ErrorBoundSampler

    sc, O = env["sc"], env["O"]
    nid = "object"
    a = env["art"][nid]
    dirs, cam = O.camera_rays(sc.uv, sc.extrinsics, sc.intrinsics)
    P = dirs.shape[1]
    dirs = dirs.reshape(-1, 3).contiguous()
    cam = cam.unsqueeze(1).repeat(1, P, 1).reshape(-1, 3).contiguous()
    g = torch.Generator().manual_seed(3)
    cfg = sc.sampler
    rand = dict(jitter=torch.rand(P, cfg["N_samples_eval"], generator=g), u=torch.rand(P, cfg["N_samples"], generator=g),
                extra_idx=torch.randperm(cfg["N_samples_eval"], generator=g)[: cfg["N_samples_extra"]])

    def q(pts):
        return O.sdf_mlp(O.rigid_inverse_warp(pts, a["tfs"][0]), sc.sdf_state[nid])[:, 0]

    with torch.no_grad():
        zo, it_o = O.error_bound_sample(q, dirs, cam, O.density_beta(sc.beta[nid]), cfg, sc.bounding_sphere, rand=rand)
    node = env["net"].nodes[nid]
    pose, keep, _, _ = node.articulate(env["inp"])
    dev = env["dev"]
    # extras index into the FINAL z buffer (n = rounds * N_eval entries): valid for any round count only below N_eval
    rnd = dict(jitter=rand["jitter"].to(dev).contiguous(), u=rand["u"].to(dev).contiguous(),
               extra_idx=rand["extra_idx"].to(torch.int32).to(dev).contiguous())
    z, iters = ErrorBoundSampler(node).get_z_vals(dirs.to(dev), cam.to(dev), pose, 1, rand=rnd)
    ctx.check()
    assert int(iters.item()) == it_o
    z = z.cpu()
    assert (z[:, 1:] >= z[:, :-1]).all()
    frac = ((z - zo).abs() <= 1e-4 * 8.0).float().mean().item()
    assert frac >= 0.93, f"only {frac:.3f} of the training-mode z_vals within tolerance"


def test_chunked_render_reproduces_the_reference_chunk_semantics(env, ctx):
    """forward_fg(chunk=n): the sampler's convergence flag is global over ONE call (ray_sampler.py:244), so rendering in the
    reference's pixel chunks (eval_datasets.py:13) gives each chunk its own round count; the result must equal the oracle run with
    the same chunking, and the chunk calls must be exactly what separate calls produce."""
    from hold_b200 import scene_io

    sc, O, dev = env["sc"], env["O"], env["dev"]
    inp = scene_io.scene_input(sc, dev)
    out = env["net"].forward_fg(inp, chunk=16)
    ctx.check()
    R = sc.H * sc.W
    assert out["fg_rgb"].shape[0] == R and out["sampler_iters_per_chunk"].shape[0] == R // 16
    ref, _ = O.render_scene(sc, chunk=16, stable_ties=True)
    for c in (0, len(ref) - 1):
        sub = scene_io.scene_input(sc, dev, ray_ids=torch.arange(c * 16, (c + 1) * 16))
        alone = env["net"].forward_fg(sub)
        assert torch.equal(alone["fg_rgb"], out["fg_rgb"][c * 16:(c + 1) * 16])
        for k, nid in enumerate(sc.node_ids):
            assert int(out["sampler_iters_per_chunk"][c][k]) == ref[c]["nodes"][k]["iters"], f"chunk {c} {nid}: round count"
            a, b = out[f"{nid}.fg_rgb"][c * 16:(c + 1) * 16].cpu(), ref[c]["render"][k]["fg_rgb"]
            assert (a.reshape(b.shape) - b).abs().max().item() < 1e-3
