"""CPU-only: the cluster-pruned, seeded exact KNN-15 of the hand inverse-warp kernel (hold_b200/csrc/knn_phases.h, run here on
the host) is bit-identical to brute force in (distance, index) order — near the hand, far from it, on vertices, with duplicate
vertices (ties), on an articulated (posed) hand whose groups were formed on the canonical one, and across large steps along
the ray (fallback path)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    out = os.path.join(ROOT, "tests", "_build", "libknn_host.so")
    src = os.path.join(ROOT, "tests", "host", "knn_host.cpp")
    hdr = os.path.join(ROOT, "hold_b200", "csrc", "knn_phases.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def test_clustered_knn_is_exact():
    from hold_b200 import synth

    lib = _lib()
    import torch
    from oracle import hold_oracle as O

    m = synth.make_mano_struct(0)
    cano = np.ascontiguousarray(m["v_template"].numpy().astype(np.float32))
    skin = np.ascontiguousarray(m["lbs_weights"].numpy().astype(np.float32))
    g = torch.Generator().manual_seed(1)
    pose = torch.randn(1, 48, generator=g) * 0.5                       # a strongly articulated hand: groups spread out
    posed = O.mano_server(m, torch.ones(1) * 4.7, torch.tensor([[0.3, -0.2, 0.1]]), pose, torch.zeros(1, 10))["verts"][0]
    verts = np.ascontiguousarray(posed.numpy().astype(np.float32))
    verts[100] = verts[7]                       # duplicate vertices: exact distance ties, resolved by index
    verts[650] = verts[7]
    rng = np.random.default_rng(0)
    n_rays, ns = 48, 96
    cam = (rng.normal(size=(n_rays, 3)) * 2.0 + np.array([0, 0, -4.0])).astype(np.float32)
    target = verts[rng.integers(0, 778, n_rays)] + rng.normal(size=(n_rays, 3)).astype(np.float32) * 0.05 * np.ptp(verts)
    dirs = target - cam
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    z = np.sort(rng.uniform(0.0, 9.0, size=(n_rays, ns)).astype(np.float32), axis=1)
    z[:, ns // 2:] = np.sort(np.linalg.norm(target - cam, axis=1, keepdims=True) + rng.normal(size=(n_rays, ns - ns // 2)).astype(np.float32) * 0.02, axis=1)
    z[0, 5] = 0.0                                # the camera itself
    cam[1] = verts[7]; z[1, 0] = 0.0             # a sample exactly on a (triplicated) vertex
    idx = np.zeros((n_rays, ns, 15), np.int32)
    dist = np.zeros((n_rays, ns, 15), np.float32)
    fb, vis = C.c_int(0), C.c_double(0.0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.knn_cluster_host(C.c_int(n_rays), C.c_int(ns), vp(verts), vp(cano), vp(skin), vp(cam), vp(dirs), vp(z), vp(idx), vp(dist),
                                C.byref(fb), C.byref(vis)) == 0
    print(f"groups visited per sample: {vis.value:.1f} of 49; full-scan fallbacks: {fb.value} of {n_rays * (ns - 1)}")
    assert vis.value < 25.0, "the pruning must prune"
    # brute force with the reference's float32 expression, points formed as cam + z * dir in float32 (two ops)
    pts = (cam[:, None, :] + (z[:, :, None] * dirs[:, None, :]).astype(np.float32)).astype(np.float32)
    diff = (pts[:, :, None, :] - verts[None, None, :, :]).astype(np.float32)
    sq = (diff * diff).astype(np.float32)
    d = ((sq[..., 0] + sq[..., 1]).astype(np.float32) + sq[..., 2]).astype(np.float32)        # [R, S, 778]
    order = np.lexsort((np.broadcast_to(np.arange(778), d.shape), d), axis=-1)[..., :15]
    assert np.array_equal(idx, order.astype(np.int32))
    assert np.array_equal(dist, np.take_along_axis(d, order, -1))
