"""Marching cubes on the GPU (hold_mc_mark / hold_mc_emit, hold_b200.meshing.marching_cubes) against its numpy restatement
(oracle/marching_cubes.py; pinned by tests/test_cpu_mc.py to geometric properties and, through mc_phases.h compiled on the host, to
the kernels' own code): vertices bit for bit, faces exactly; generate_mesh(backend="gpu") on an analytic SDF gives one closed,
outward-oriented component of the right volume."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _closed_oriented(faces):
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0).astype(np.int64)
    m = faces.max() + 1
    key, rev = e[:, 0] * m + e[:, 1], e[:, 1] * m + e[:, 0]
    return np.unique(key).size == key.size and np.array_equal(np.sort(key), np.sort(rev))


def test_marching_cubes_equals_restatement(ctx):
    from hold_b200 import meshing
    from oracle.marching_cubes import marching_cubes

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(2)
    a = np.linspace(-1, 1, 65, dtype=np.float32)
    x, y, z = np.meshgrid(a, a, a, indexing="ij")
    cases = [(np.sqrt(x * x + y * y + z * z) - 0.71, 0.0), ((np.sqrt(x * x + y * y) - 0.6) ** 2 + z * z - 0.05, 0.0),
             (rng.standard_normal((17, 33, 9)).astype(np.float32), 0.15), (np.round(rng.standard_normal((12, 12, 12)) * 2).astype(np.float32) / 2, 0.5),
             (np.ones((6, 7, 8), np.float32), 0.0)]
    for vol, level in cases:
        v, f = meshing.marching_cubes(ctx, torch.as_tensor(vol, device=dev), level)
        ctx.check()
        v0, f0 = marching_cubes(vol, level)
        assert tuple(v.shape) == v0.shape and tuple(f.shape) == f0.shape, (v.shape, v0.shape, f.shape, f0.shape)
        assert np.array_equal(v.cpu().numpy().view(np.uint32), v0.view(np.uint32))
        assert np.array_equal(f.cpu().numpy(), f0)


def test_generate_mesh_gpu_backend(ctx):
    """generate_mesh(backend="gpu") = MISE value grid + GPU marching cubes + largest component, on an analytic SDF whose surface lies
    inside the bounding box the reference derives from the canonical vertices (utils/meshing.py:10-24): a closed, outward-oriented
    sphere of the right volume with its vertices on the level set."""
    from hold_b200 import meshing

    dev = torch.device("cuda", 0)
    r = 0.3
    g = torch.Generator().manual_seed(0)
    on_sphere = torch.nn.functional.normalize(torch.randn(2000, 3, generator=g), dim=1) * r + torch.tensor([0.05, -0.02, 0.01])
    centre = torch.tensor([0.05, -0.02, 0.01], device=dev)
    func = lambda pts: (pts - centre).norm(dim=1) - r
    v, f = meshing.generate_mesh(ctx, func, on_sphere.numpy(), res_init=16, res_up=2, backend="gpu")
    ctx.check()
    assert f.shape[0] > 1000 and _closed_oriented(f)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol = np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0
    assert abs(vol - 4 / 3 * np.pi * r ** 3) <= 0.02 * 4 / 3 * np.pi * r ** 3, vol          # > 0: outward normals
    assert np.abs(np.linalg.norm(v - centre.cpu().numpy(), axis=1) - r).max() <= 2e-4
