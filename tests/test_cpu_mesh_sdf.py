"""CPU-only, SURVEY §8f rank 3: the per-point arithmetic of k_mesh_sdf (hold_b200/csrc/mesh_sdf_phases.h, compiled for the
host) against the float64 oracle written with different formulas (oracle/mesh_sdf_oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import mesh_sdf_oracle as MO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    out = os.path.join(ROOT, "tests", "_build", "libmesh_sdf_host.so")
    src = os.path.join(ROOT, "tests", "host", "mesh_sdf_host.cpp")
    hdr = os.path.join(ROOT, "hold_b200", "csrc", "mesh_sdf_phases.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def test_mesh_sdf_kernel_arithmetic_on_host():
    lib = _lib()
    verts, faces = MO.star_mesh(level=3, seed=1)
    assert faces.shape == (512, 3)
    rng = np.random.default_rng(0)
    P = 3000
    pts = (rng.uniform(-1.5, 1.5, size=(P, 3))).astype(np.float32)
    # a few exactly-on-vertex / on-face points: distance must be 0 (sign is undefined there)
    pts[:5] = verts[:5]
    pts[5] = verts[faces[7]].mean(0)
    ref, _ = MO.signed_distance(pts, verts, faces)
    sdf = np.full(P, np.nan, np.float32)
    fidx = np.full(P, -7, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.mesh_sdf_host(C.c_int(P), C.c_int(len(verts)), C.c_int(len(faces)), vp(pts), vp(verts), vp(faces), vp(sdf), vp(fidx)) == 0
    assert np.isfinite(sdf).all() and (fidx >= 0).all() and (fidx < len(faces)).all()
    assert np.abs(np.abs(sdf) - np.abs(ref)).max() < 2e-6, np.abs(np.abs(sdf) - np.abs(ref)).max()
    off = np.abs(ref) > 1e-4                       # away from the surface the sign is well defined
    assert (np.sign(sdf[off]) == np.sign(ref[off])).all()
    assert (np.abs(sdf[:6]) < 1e-6).all()
    assert 0.1 < (ref[off] < 0).mean() < 0.3       # the sample really has inside and outside points


def test_off_in_surface_reduction_semantics():
    """check_off_in_surface_points_cano_mesh (volsdf_utils.py:209-217) in numpy: what k_off_in_surface computes."""
    rng = np.random.default_rng(2)
    sd = rng.normal(size=(40, 9)).astype(np.float32)
    m = sd.min(1)
    off, inn = m > 0.05, m <= 0.0
    assert off.dtype == bool and not (off & inn).any()
