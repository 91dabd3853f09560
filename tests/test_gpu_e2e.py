"""End-to-end parity of hold_render_fg (through the host mirror) against the committed golden fixtures,
which were produced by the REFERENCE's own modules (oracle/ref_harness.py golden)."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def e2e_close(a, b, name, tol=1e-4, tol_max=1e-2, frac_min=0.97):
    """End-to-end criteria sit at the reference's own noise floor (tools/noise_floor.py: with exp() perturbed by
    +-1 ulp the reference's per-node pixels move by up to 4e-3 with ~1 % beyond 1e-4, composite pixels by up to
    6e-3 with 8-27 % beyond 1e-4 because merge_factors interleaves the nodes' noisy sample positions)."""
    a, b = a.detach().float().cpu().reshape(b.shape), b.float()
    scale = max(1.0, b.abs().max().item())
    d = (a - b).abs()
    frac = (d <= tol * scale).float().mean().item()
    print(f"  e2e {name}: within {tol:.0e}: {frac:.4f}, max {d.max().item() / scale:.2e} of scale")
    assert frac >= frac_min, f"{name}: only {frac:.4f} within {tol} (max {d.max().item():.3e})"
    assert d.max().item() <= tol_max * scale, f"{name}: max|d| {d.max().item():.3e}"


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-3] for p in GOLD])
@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_golden(path, mode, ctx):
    from hold_b200 import capi, scene_io, synth

    if mode == "tc" and not getattr(capi, "TC_READY", False):
        pytest.skip("tcgen05 MLP path not enabled in this build")
    rec = torch.load(path)
    sc = synth.make_scene(**rec["scene_kwargs"])
    for nid in sc.node_ids:
        sc.beta[nid] = torch.tensor(rec["beta"])
    net = scene_io.build_net(sc, ctx, capi.MLP_TC if mode == "tc" else capi.MLP_FP32)
    dev = torch.device("cuda", 0)
    out = net.forward_fg(scene_io.scene_input(sc, dev, ray_ids=rec["ray_ids"]))
    ctx.check()
    for nid in sc.node_ids:
        art = out["articulation"][nid]
        e2e_close(art["verts"], rec["art"][nid]["verts"], f"{nid}.verts", 1e-5, 1e-4)
        if nid != "object":
            e2e_close(art["tfs"], rec["art"][nid]["tfs"], f"{nid}.tfs", 1e-5, 1e-4)
            e2e_close(art["jnts"], rec["art"][nid]["jnts"], f"{nid}.jnts", 1e-5, 1e-4)
        # Both arithmetic modes are held to the same bar (round 2: accumulator-truncation compensation brought the tensor-core
        # sdf to fp32 level, profiles/r02_tc_accumulator_bias.md): >= 95 % of the pixels within 1e-4 and max <= 1e-2, i.e. <= 3x
        # the worst values measured over the three goldens (fp32: 97.9 % / 3.0e-3; tensor cores: 96.1 % / 3.4e-3).  The residual is
        # the sampler's sensitivity to the last bit of exp() (tests/test_gpu_sampler_rounds.py quantifies it round by round).
        fmin, emax = 0.95, 1e-2
        for k in ("fg_rgb", "mask_prob", "depth", "normal", "bg_weights"):
            e2e_close(out[f"{nid}.{k}"], rec["render"][nid][k], f"{nid}.{k}", tol_max=emax, frac_min=fmin)
    # Composite: the reference sorts the concatenated z of all nodes with an UNSTABLE torch.sort; exact z ties
    # between nodes are the norm (~18 per ray: shared uniform grid, near, far), and their order alone moves the
    # reference's composite by up to 6e-2 (depth) on ~15 % of the pixels ("ties").  hold_b200
    # implements the stable order; the check is against the golden PER-NODE factors re-composited in that order.
    from oracle import hold_oracle as O

    fl = []
    for nid in sc.node_ids:
        n = rec["nodes"][nid]
        sem = torch.zeros(n["z_vals"].shape[0], n["z_vals"].shape[1], 4)
        sem[:, :, O.CLASS_ID[nid]] = 1.0
        fl.append(dict(color=n["color"], normal=n["normal"], density=n["density"], semantics=sem, z_vals=n["z_vals"]))
    assert (O.composite(fl)["comp"]["fg_rgb"] - rec["render"]["comp"]["fg_rgb"]).abs().max() < 1e-6  # oracle == reference
    canon = O.composite(fl, stable=True)["comp"]
    # tools/noise_floor.py: with exp() perturbed by +-1 ulp the REFERENCE's own 3-node composite moves by up to
    # 1e-2 (rgb) / 3.7e-2 (depth) with 76-95 % of the pixels beyond 1e-4 while its per-node renders stay within
    # 1.3e-4 — the interleaving of the nodes' sample sets amplifies sample-position noise.  The composite is
    # therefore held to a mean/max bound here; its 1e-5 stage parity (same factors in) is test_gpu_stages.py.
    # Bounds per golden = <= 3x the values measured on hardware in round 2 (worst of both modes and of the six quantities):
    #   c1 64x64 S=32 n=2: mean 3.0e-5, max 1.5e-2, 98.96 % within 1e-4;  S=128 n=2 beta 0.03: 1.15e-3 / 4.9e-2 / 77 %;
    #   S=128 n=3 beta 0.05: 1.34e-3 / 5.8e-2 / 66 %.
    name = os.path.basename(path)[:-3]
    mean_max, abs_max, frac_min = {"c1_64x64_S32_n2": (1e-4, 4.5e-2, 0.97), "s128_n2_beta03": (3.5e-3, 1.5e-1, 0.60),
                                   "s128_n3_beta05": (4e-3, 1.7e-1, 0.50)}.get(name, (4e-3, 1.7e-1, 0.50))
    for k in ("fg_rgb", "mask_prob", "depth", "normal", "fg_semantics", "bg_weights"):
        a, b = out[k].detach().float().cpu().reshape(canon[k].shape), canon[k]
        d = (a - b).abs()
        frac = (d <= 1e-4 * max(1.0, b.abs().max().item())).float().mean().item()
        print(f"  e2e comp.{k}: mean {d.mean().item():.2e} max {d.max().item():.2e} within 1e-4: {frac:.4f}")
        assert d.mean().item() <= mean_max and d.max().item() <= abs_max and frac >= frac_min, \
            f"comp.{k}: mean {d.mean().item():.2e} max {d.max().item():.2e} within-1e-4 {frac:.3f}"
