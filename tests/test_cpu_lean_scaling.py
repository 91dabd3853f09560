"""CPU-only: the operand scaling of the tcgen05 SDF kernels, emulated in numpy — fp16 hi/lo split of A and W, three products
per term, fp32 accumulation — for the default scheme (A = 64 softplus, W * 2^10, accumulator * 2^-16) and the LEAN scheme
(HOLD_TC_LEAN=1: A = S(t), activation-fed weight columns * ln2/100 * 2^17, embedding-fed columns * 2^11, accumulator =
2^17 z, t = acc * (100 log2 e / 2^17) + 100 log2 e * b, head row * ln2/100).  Both must reproduce the fp64 network to the
split-precision level; this pins the constants of mlp_tc.cuh (kLean*) and the per-column split of k_tc_pack."""
import numpy as np
import torch

from hold_b200 import synth
from oracle import hold_oracle as O

LOG2E100 = 144.26950408889634
ACT = 0.6931471805599453 * 0.01


def _split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def _mm3(a, w):
    """[P,K] x [N,K]^T with hi/lo operands: hi*hi + lo*hi + hi*lo, fp32 accumulate (order-insensitive emulation)."""
    ah, al = _split(a.astype(np.float32))
    wh, wl = _split(w.astype(np.float32))
    return (ah @ wh.T + al @ wh.T + ah @ wl.T).astype(np.float32)


def _folded(sd, kind):
    """weight-norm folded, hand pose columns dropped, skip layer pre-scaled: what k_tc_pack sees (api.cu set_weights)."""
    Ws, bs = [], []
    for l in range(9):
        v, g, b = sd[f"lin{l}.weight_v"].double().numpy(), sd[f"lin{l}.weight_g"].double().numpy(), sd[f"lin{l}.bias"].double().numpy()
        w = v * (g / np.linalg.norm(v, axis=1, keepdims=True))
        if l == 0:
            w = w[:, :39]
        if l == 4:
            w = w / np.sqrt(2.0)
        Ws.append(w), bs.append(b)
    return Ws, bs


def _embed(x):
    out = [x]
    for k in range(6):
        out += [np.sin(x * 2.0**k), np.cos(x * 2.0**k)]
    return np.concatenate(out, -1)


def _run(Ws, bs, x, scheme):
    emb = _embed(x).astype(np.float32)                       # [P,39]
    a = 64.0 * emb                                           # layer-0 operand in both schemes
    for l in range(8):
        W = Ws[l]
        if scheme == "default":
            acc = _mm3(a, 1024.0 * W)
            z = acc * np.float32(2.0**-16) + bs[l].astype(np.float32)
            h = 64.0 * np.logaddexp(0.0, 100.0 * z.astype(np.float64)).astype(np.float32) / 100.0
        else:
            cs = np.full(W.shape[1], ACT * 2.0**17)
            if l == 0:
                cs[:] = 2.0**11
            if l == 4:
                cs[217:] = 2.0**11
            acc = _mm3(a, W * cs[None, :])
            t = acc * np.float32(LOG2E100 / 2.0**17) + (LOG2E100 * bs[l]).astype(np.float32)
            h = (np.maximum(t, 0) + np.log2(1.0 + np.exp2(-np.abs(t.astype(np.float64))))).astype(np.float32)   # S(t)
        if l == 3:
            h = np.concatenate([h[:, :217], 64.0 * emb], 1)   # skip connection: embedding columns carry the 2^6 scale
        a = h
    w_sdf, b_sdf = Ws[8][0], bs[8][0]
    if scheme == "default":
        return (a.astype(np.float64) @ w_sdf) / 64.0 + b_sdf
    return a.astype(np.float64) @ (w_sdf * ACT) + b_sdf


def test_operand_scaling_schemes_match_fp64():
    sd = synth.make_sdf_state("hand", 0, 0.45)
    Ws, bs = _folded(sd, "hand")
    g = torch.Generator().manual_seed(1)
    x = ((torch.rand(400, 3, generator=g) - 0.5) * 1.6).double().numpy()
    ref = O.sdf_mlp(torch.from_numpy(x).float(), None, sd, "hand") if False else None
    # fp64 network
    h = _embed(x)
    for l in range(8):
        z = h @ Ws[l].T + bs[l]
        h = np.logaddexp(0.0, 100.0 * z) / 100.0
        if l == 3:
            h = np.concatenate([h[:, :217], _embed(x)], 1)
    exact = h @ Ws[8][0] + bs[8][0]
    for scheme in ("default", "lean"):
        got = _run(Ws, bs, x, scheme)
        err = np.abs(got - exact).max() / max(1.0, np.abs(exact).max())
        assert err < 3e-6, f"{scheme}: {err:.2e}"
