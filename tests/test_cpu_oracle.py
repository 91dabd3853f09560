"""CPU-only: the oracle (oracle/hold_oracle.py) against the committed golden fixtures, which were produced by the
REFERENCE's own modules (oracle/ref_harness.py golden)."""
import glob
import os

import pytest
import torch

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def _close(a, b, tol=1e-4, tol_max=3e-3, w=None, frac_min=0.97):
    a, b = a.float(), b.float()
    d = (a - b).abs()
    scale = max(1.0, b.abs().max().item())
    frac = (d <= tol * scale).float().mean().item()
    if w is not None:
        m = w > 1e-4
        while m.dim() < d.dim():
            m = m.unsqueeze(-1)
        d = d * m
    return frac >= frac_min and d.max().item() <= tol_max * scale


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-3] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    from hold_b200 import synth
    from oracle import hold_oracle as O

    torch.set_num_threads(min(8, os.cpu_count() or 1))
    rec = torch.load(path)
    sc = synth.make_scene(**rec["scene_kwargs"])
    for nid in sc.node_ids:
        sc.beta[nid] = torch.tensor(rec["beta"])
    ids = rec["ray_ids"][:48]  # a subset keeps the CPU suite short; the sampler flag is per call, so compare per-ray data loosely
    outs, art = O.render_scene(sc, ray_ids=rec["ray_ids"])
    for k, nid in enumerate(sc.node_ids):
        assert torch.equal(art[nid]["tfs"], rec["art"][nid]["tfs"]) or _close(art[nid]["tfs"], rec["art"][nid]["tfs"], 1e-6, 1e-5)
        assert _close(art[nid]["verts"], rec["art"][nid]["verts"], 1e-6, 1e-5)
        n, g = outs[0]["nodes"][k], rec["nodes"][nid]
        w = outs[0]["render"][k]["fg_weights"]
        # Bounds = <= 3x what this comparison measures (oracle here vs goldens written by the reference modules in another process:
        # same algorithm, same libm, different GEMM blocking).  Measured worst case over the three goldens: per-sample tensors
        # 97.4 % within 1e-4 (normals, weight-carrying max 8.3e-2; others <= 5.7e-3); per-node renders 95.05 % / max 3.1e-3;
        # composite mean 4.0e-4, max 2.7e-2.  That the SAME algorithm moves this much between two processes is the reference's own
        # sensitivity (tests/test_gpu_sampler_rounds.py); the tight same-process pin is tests/test_cpu_oracle_pin.py.
        for key in ("z_vals", "sdf", "canonical_pts", "normal", "color"):
            assert _close(n[key], g[key], w=w, tol_max=(1e-1 if key == "normal" else 1.5e-2), frac_min=0.97), f"{nid}.{key}"
        for key in ("fg_rgb", "mask_prob", "depth", "normal", "bg_weights"):
            assert _close(outs[0]["render"][k][key], rec["render"][nid][key], tol_max=1e-2, frac_min=0.94), f"{nid}.render.{key}"
    for key in ("fg_rgb", "mask_prob", "depth", "normal", "fg_semantics", "bg_weights"):
        d = (outs[0]["render"]["comp"][key] - rec["render"]["comp"][key]).abs()
        assert d.mean().item() <= 1.2e-3 and d.max().item() <= 8e-2, f"comp.{key}"


def test_oracle_edge_cases():
    from oracle import hold_oracle as O

    # ray that misses the bounding sphere -> error, not exit() (engine/ray_sampler.py:15-18)
    with pytest.raises(O.RayMissesSphere):
        O.sphere_far(torch.tensor([[10.0, 0.0, 0.0]]), torch.tensor([[0.0, 1.0, 0.0]]), 1.0)
    # merge_factors drops (n-1) head / n tail samples (hold_utils.py:115-119)
    R, S = 3, 5
    fl = []
    for k in range(3):
        z = torch.sort(torch.rand(R, S), 1).values
        fl.append(dict(color=torch.rand(R, S, 3), normal=torch.rand(R, S, 3), density=torch.rand(R, S, 1),
                       semantics=torch.zeros(R, S, 4), z_vals=z))
    m = O.merge_factors(fl)
    assert m["z_vals"].shape == (R, 3 * S - 2 * 3 + 1) and (m["z_vals"][:, 1:] >= m["z_vals"][:, :-1]).all()
    # Laplace density is 1/(2 beta) at the surface and monotone in -sdf
    s = torch.linspace(-1, 1, 11)
    d = O.laplace_density(s, torch.tensor(0.1))
    assert abs(d[5].item() - 5.0) < 1e-6 and (d[:-1] >= d[1:]).all()
    # knn contract: ascending squared distances, K smallest
    p, v = torch.rand(7, 3), torch.rand(50, 3)
    dd, ii = O.knn_points(p, v, 15)
    assert (dd[:, 1:] >= dd[:, :-1]).all() and torch.allclose(dd, ((p[:, None] - v[ii]) ** 2).sum(-1))


def test_background_oracle_matches_reference_golden():
    """oracle.background against outputs of the reference's Background class
    (oracle/ref_harness.py golden_background)."""
    import torch
    from hold_b200 import synth
    from oracle import hold_oracle as O

    rec = torch.load(os.path.join(os.path.dirname(__file__), "golden", "background", "bg_8x8_B2.pt"))
    i, ref = rec["in"], rec["out"]
    sdf_sd, rgb_sd = synth.make_bg_state(i["bg_state_seed"])
    o = O.background(i["bg_weights"], i["ray_dirs"], i["cam_loc"], i["frame_code"], i["frame_of_ray"], sdf_sd, rgb_sd, i["r_sphere"])
    for got, key in zip(o, ("bg_rgb", "bg_rgb_only", "bg_semantics", "bg_z_vals")):
        assert torch.isfinite(ref[key]).all()
        assert (got - ref[key]).abs().max().item() <= 2e-6, key
