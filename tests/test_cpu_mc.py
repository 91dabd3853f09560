"""CPU-only: the marching-cubes tables (tools/gen_mc_tables.py, derived case by case, not copied) and the numpy restatement the GPU
kernels are held to (oracle/marching_cubes.py): the committed tables are what the generator produces; meshes of analytic fields are
closed 2-manifolds with consistent outward orientation (every directed edge is matched by its reverse exactly once), their vertices
lie on the level set, their volume and area converge to the analytic values; random noise (every ambiguous face configuration)
stays watertight."""
import importlib.util
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_mc_tables", os.path.join(ROOT, "tools", "gen_mc_tables.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _grid(n, lo=-1.0, hi=1.0):
    a = np.linspace(lo, hi, n, dtype=np.float32)
    return np.meshgrid(a, a, a, indexing="ij"), (hi - lo) / (n - 1)


def _check_closed_oriented(faces):
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0).astype(np.int64)
    key = e[:, 0] * (faces.max() + 1) + e[:, 1]
    rev = e[:, 1] * (faces.max() + 1) + e[:, 0]
    assert np.unique(key).size == key.size, "a directed edge is used twice (inconsistent orientation or non-manifold)"
    assert np.array_equal(np.sort(key), np.sort(rev)), "an edge has no oppositely directed partner (hole)"


def _volume_area(v, f):
    a, b, c = v[f[:, 0]].astype(np.float64), v[f[:, 1]].astype(np.float64), v[f[:, 2]].astype(np.float64)
    vol = np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()
    return vol, area


def test_committed_tables_are_the_generated_ones():
    g = _gen()
    tri, width, _ = g.tables()
    d = json.load(open(os.path.join(ROOT, "oracle", "mc_tables.json")))
    assert d["width"] == width == 5
    assert d["ntri"] == [len(t) for t in tri] and sum(d["ntri"]) == 820
    for c in range(256):
        assert d["tri"][c][: 3 * len(tri[c])] == [e for t in tri[c] for e in t]
    hdr = open(os.path.join(ROOT, "hold_b200", "csrc", "mc_tables.h")).read()
    assert "{" + ", ".join(map(str, d["tri"][105])) + "}" in hdr
    # complementary cases cut the same edges (the surface separates the same corner sets)
    for c in range(256):
        assert {e for t in tri[c] for e in t} == {e for t in tri[255 - c] for e in t}


def test_sphere_and_torus():
    from oracle.marching_cubes import marching_cubes

    for n in (24, 48):
        (x, y, z), h = _grid(n)
        r = 0.7
        v, f = marching_cubes(np.sqrt(x * x + y * y + z * z) - r, 0.0)
        _check_closed_oriented(f)
        vol, area = _volume_area(v * h, f)
        assert vol > 0, "normals must point outwards (towards increasing values)"
        assert abs(vol - 4 / 3 * np.pi * r ** 3) <= 0.03 * (24 / n) ** 2 * 4 / 3 * np.pi * r ** 3
        assert abs(area - 4 * np.pi * r * r) <= 0.03 * (24 / n) ** 2 * 4 * np.pi * r * r
        p = v * h - 1.0
        assert np.abs(np.linalg.norm(p, axis=1) - r).max() <= 0.6 * h * h / r + 1e-6      # linear interpolation error of a curved field
    (x, y, z), h = _grid(40)
    R, r = 0.6, 0.22
    v, f = marching_cubes((np.sqrt(x * x + y * y) - R) ** 2 + z * z - r * r, 0.0)
    _check_closed_oriented(f)
    vol, _ = _volume_area(v * h, f)
    assert abs(vol - 2 * np.pi ** 2 * R * r * r) <= 0.04 * 2 * np.pi ** 2 * R * r * r
    assert v.shape[0] - (3 * f.shape[0] // 2) + f.shape[0] == 0                             # Euler characteristic of a torus


def test_noise_is_watertight():
    from oracle.marching_cubes import marching_cubes

    rng = np.random.default_rng(0)
    vol = rng.standard_normal((14, 15, 16)).astype(np.float32)
    vol[0], vol[-1], vol[:, 0], vol[:, -1], vol[:, :, 0], vol[:, :, -1] = 3, 3, 3, 3, 3, 3   # closed box: no surface leaves the grid
    v, f = marching_cubes(vol, 0.1)
    assert f.shape[0] > 2000
    _check_closed_oriented(f)
    vol_in, _ = _volume_area(v, f)
    assert vol_in > 0
    # level exactly on grid values: such nodes count as outside, vertices may coincide with nodes, the mesh stays closed
    q = np.round(vol * 2) / 2
    v, f = marching_cubes(q, 0.5)
    _check_closed_oriented(f)
    assert np.isfinite(v).all()


def _host_lib():
    import ctypes as C
    import subprocess

    out = os.path.join(ROOT, "tests", "_build", "libmc_host.so")
    src = os.path.join(ROOT, "tests", "host", "mc_host.cpp")
    deps = [src, os.path.join(ROOT, "hold_b200", "csrc", "mc_phases.h"), os.path.join(ROOT, "hold_b200", "csrc", "mc_tables.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def host_marching_cubes(vol, level):
    """the kernels' code (mc_phases.h) run on the host + the same two exclusive scans the Python wrapper does with torch.cumsum"""
    import ctypes as C

    lib = _host_lib()
    vol = np.ascontiguousarray(vol, np.float32)
    n0, n1, n2 = vol.shape
    flags = np.zeros(n0 * n1 * n2 * 3, np.int32)
    ntri = np.zeros((n0 - 1) * (n1 - 1) * (n2 - 1), np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.mc_mark(n0, n1, n2, P(vol), C.c_float(level), P(flags), P(ntri))
    vid = (np.cumsum(flags, dtype=np.int64) - flags).astype(np.int64)
    toff = (np.cumsum(ntri, dtype=np.int64) - ntri).astype(np.int64)
    verts = np.zeros((int(flags.sum()), 3), np.float32)
    faces = np.zeros((int(ntri.sum()), 3), np.int32)
    lib.mc_emit(n0, n1, n2, P(vol), C.c_float(level), P(flags), P(vid), P(toff), P(verts), P(faces))
    return verts, faces


def test_kernel_code_on_the_host_equals_the_restatement():
    from oracle.marching_cubes import marching_cubes

    rng = np.random.default_rng(1)
    (x, y, z), _ = _grid(20)
    cases = [(np.sqrt(x * x + y * y + z * z) - 0.63, 0.0), (rng.standard_normal((9, 12, 7)).astype(np.float32), 0.2),
             (np.round(rng.standard_normal((8, 8, 8)) * 2).astype(np.float32) / 2, 0.5), (np.ones((5, 5, 5), np.float32), 0.0)]
    for vol, level in cases:
        v0, f0 = marching_cubes(vol, level)
        v1, f1 = host_marching_cubes(vol, level)
        assert v0.shape == v1.shape and f0.shape == f1.shape
        assert np.array_equal(v0.view(np.uint32), v1.view(np.uint32)), "vertex positions differ in the last bit"
        assert np.array_equal(f0, f1)
