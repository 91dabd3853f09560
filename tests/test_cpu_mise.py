"""CPU-only, SURVEY §8f rank 4: the MISE restatement the CUDA kernels run (hold_b200/csrc/mise_phases.h, compiled for the
host) against the REFERENCE's own compiled Cython MISE (oracle/_ref/mise*.so from oracle/build_ref_mise.py; the built module
travels with the repo snapshot, /root/reference is only needed to build it) — round by round and bit for bit."""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host():
    out = os.path.join(ROOT, "tests", "_build", "libmise_host.so")
    src = os.path.join(ROOT, "tests", "host", "mise_host.cpp")
    hdr = os.path.join(ROOT, "hold_b200", "csrc", "mise_phases.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.mise_host_create.restype = C.c_void_p
    lib.mise_host_create.argtypes = [C.c_int, C.c_int, C.c_float]
    lib.mise_host_query.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.mise_host_update.argtypes = [C.c_void_p, C.c_void_p]
    lib.mise_host_to_dense.argtypes = [C.c_void_p, C.c_void_p]
    lib.mise_host_destroy.argtypes = [C.c_void_p]
    return lib


def _ref_mise():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref_mise

    so = build_ref_mise.build() or next(iter(glob.glob(os.path.join(ROOT, "oracle", "_ref", "mise*.so"))), None)
    if so is None:
        pytest.skip("reference MISE not built (needs /root/reference once)")
    sys.path.insert(0, os.path.dirname(so))
    import mise

    return mise


def _field(seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.3, 0.7, size=(3, 3))
    r = rng.uniform(0.12, 0.22, size=3)

    def f(pts, R):   # union of three blobs minus a ripple, exact zeros included through quantisation
        x = pts.astype(np.float32) / np.float32(R)
        d = np.min(np.stack([np.linalg.norm(x - c[i].astype(np.float32), axis=1) - np.float32(r[i]) for i in range(3)]), 0)
        v = (d + np.float32(0.02) * np.sin(np.float32(40.0) * x[:, 0])).astype(np.float32)
        return np.round(v * 64) / 64 if seed % 2 else v     # odd seeds: many values exactly on the threshold
    return f


@pytest.mark.parametrize("res0,depth,seed", [(4, 2, 0), (4, 3, 1), (8, 2, 2), (6, 3, 3), (32, 2, 4)])
def test_mise_restatement_matches_reference_mise(res0, depth, seed):
    lib, mise = _host(), _ref_mise()
    f = _field(seed)
    ref = mise.MISE(res0, depth, 0.0)
    h = lib.mise_host_create(res0, depth, C.c_float(0.0))
    R = res0 << depth
    G = R + 1
    rounds = 0
    while True:
        pr = ref.query()
        buf = np.zeros((G**3, 3), np.int32)
        n = lib.mise_host_query(h, buf.ctypes.data_as(C.c_void_p), C.c_int(G**3))
        assert n == pr.shape[0], f"round {rounds}: {n} vs {pr.shape[0]} points"
        if n == 0:
            break
        mine = buf[:n]
        key = lambda p: (p[:, 0].astype(np.int64) * G + p[:, 1]) * G + p[:, 2]
        assert np.array_equal(np.sort(key(mine)), np.sort(key(pr))), f"round {rounds}: different point sets"
        ref.update(pr, f(pr, R).astype(np.float64))
        vals = f(mine.astype(np.int64), R).astype(np.float32)
        lib.mise_host_update(h, vals.ctypes.data_as(C.c_void_p))
        rounds += 1
    assert rounds >= 2
    dense_ref = ref.to_dense()
    out = np.empty(G**3, np.float32)
    lib.mise_host_to_dense(h, out.ctypes.data_as(C.c_void_p))
    lib.mise_host_destroy(h)
    assert np.array_equal(out.reshape(G, G, G).astype(np.float64), dense_ref)
