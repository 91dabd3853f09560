"""hold_mesh_sdf / hold_off_in_surface (the kaolin replacements of engine/volsdf_utils.py:172-217)
against the float64 oracle.  The per-point arithmetic already runs on the CPU (tests/test_cpu_mesh_sdf.py); no hardware
run existed when this was written, hence non-strict xfail."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def impl_mesh_sdf_and_off_in_surface(ctx):
    from hold_b200 import ops
    from oracle import mesh_sdf_oracle as MO

    dev = torch.device("cuda", 0)
    verts, faces = MO.star_mesh(level=3, seed=1)
    rng = np.random.default_rng(0)
    B, R, S = 2, 60, 7
    pts = rng.uniform(-1.5, 1.5, size=(B, R * S, 3)).astype(np.float32)
    vb = np.stack([verts, verts * 1.1]).astype(np.float32)           # per-frame meshes
    ref = np.stack([MO.signed_distance(pts[b], vb[b], faces)[0] for b in range(B)])
    sd = ops.compute_mano_cano_sdf(ctx, torch.from_numpy(vb).to(dev), torch.from_numpy(faces).to(dev), torch.from_numpy(pts).to(dev))
    ctx.check()
    sd = sd.cpu().numpy()
    assert np.abs(np.abs(sd) - np.abs(ref)).max() < 2e-6
    away = np.abs(ref) > 1e-4
    assert (np.sign(sd[away]) == np.sign(ref[away])).all()
    # shared mesh (the hand's canonical mesh is the same for every frame)
    sd1 = ops.compute_mano_cano_sdf(ctx, torch.from_numpy(verts).to(dev), torch.from_numpy(faces).to(dev), torch.from_numpy(pts).to(dev)).cpu().numpy()
    assert np.abs(sd1[0] - sd[0]).max() == 0.0
    off, inn = ops.check_off_in_surface_points_cano_mesh(ctx, torch.from_numpy(vb).to(dev), torch.from_numpy(faces).to(dev),
                                                         torch.from_numpy(pts).to(dev), B * R, threshold=0.05)
    m = sd.reshape(B * R, S).min(1)
    assert (off.cpu().numpy() == (m > 0.05)).all() and (inn.cpu().numpy() == (m <= 0.0)).all()


def test_mesh_sdf_and_off_in_surface(isolated):
    isolated("tests/test_gpu_mesh_sdf.py", "impl_mesh_sdf_and_off_in_surface")
