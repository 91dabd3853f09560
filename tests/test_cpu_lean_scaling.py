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


def test_reverse_mode_chain_with_t_stash_matches_autograd():
    """The algebra of the FAST reverse-mode kernel (mlp_tc_fast.cuh: LEAN forward, stash of t, softplus' = sigmoid_t(t) evaluated
    in the backward rounds, seed 64/(ln2/100) * w_t * s_7, transposed images at the plain 2^10 / 2^6 scales, embedding columns
    chained with d embed / d x) emulated with the same split-operand products: sdf, gradient and feature against fp64 autograd."""
    sd = synth.make_sdf_state("hand", 0, 0.45)
    Ws, bs = _folded(sd, "hand")
    g = torch.Generator().manual_seed(2)
    x = ((torch.rand(200, 3, generator=g) - 0.5) * 1.6).double()
    # fp64 autograd reference
    xt = x.clone().requires_grad_()
    def emb_t(v):
        out = [v]
        for k in range(6):
            out += [torch.sin(v * 2.0**k), torch.cos(v * 2.0**k)]
        return torch.cat(out, -1)
    h = emb_t(xt)
    for l in range(8):
        z = h @ torch.from_numpy(Ws[l]).T + torch.from_numpy(bs[l])
        h = torch.nn.functional.softplus(z, beta=100)
        if l == 3:
            h = torch.cat([h[:, :217], emb_t(xt)], 1)
    out8 = h @ torch.from_numpy(Ws[8]).T + torch.from_numpy(bs[8])
    sdf_ref = out8[:, 0]
    grad_ref = torch.autograd.grad(sdf_ref.sum(), xt)[0].numpy()
    feat_ref, sdf_ref = out8[:, 1:].detach().numpy(), sdf_ref.detach().numpy()
    # ---- emulation
    xn = x.numpy()
    emb = _embed(xn).astype(np.float32)
    demb = np.zeros((xn.shape[0], 39, 3), np.float32)           # d embed_e / d x_d (embedders.py layout)
    for e in range(39):
        d = e % 3
        if e < 3:
            demb[:, e, d] = 1.0
        else:
            qq = (e - 3) // 3
            f = 2.0 ** (qq >> 1)
            demb[:, e, d] = (-f * np.sin(xn[:, d] * f)) if (qq & 1) else (f * np.cos(xn[:, d] * f))
    sigmoid_t = lambda t: np.where(t >= 0, 1.0 / (1.0 + np.exp2(-np.abs(t))), np.exp2(-np.abs(t)) / (1.0 + np.exp2(-np.abs(t)))).astype(np.float32)
    a, stash = 64.0 * emb, []
    for l in range(8):
        cs = np.full(Ws[l].shape[1], ACT * 2.0**17)
        if l == 0:
            cs[:] = 2.0**11
        if l == 4:
            cs[217:] = 2.0**11
        t = _mm3(a, Ws[l] * cs[None, :]) * np.float32(LOG2E100 / 2.0**17) + (LOG2E100 * bs[l]).astype(np.float32)
        stash.append(t)
        S = (np.maximum(t, 0) + np.log2(1.0 + np.exp2(-np.abs(t.astype(np.float64))))).astype(np.float32)
        a = np.concatenate([S[:, :217], 64.0 * emb], 1) if l == 3 else S
    sdf = a.astype(np.float64) @ (Ws[8][0] * ACT) + bs[8][0]
    feat = _mm3(a, Ws[8][1:] * (ACT * 2.0**17)) * np.float32(2.0**-17) + bs[8][1:].astype(np.float32)
    gA = (np.float32(64.0 / ACT) * (Ws[8][0] * ACT).astype(np.float32)[None, :] * sigmoid_t(stash[7])).astype(np.float32)   # 64 * g_7
    grad = np.zeros((xn.shape[0], 3), np.float64)
    for l in range(7, -1, -1):
        acc = _mm3(gA, 1024.0 * Ws[l].T) * np.float32(2.0**-16)          # g_l . W_l over the transposed image
        if l == 4:
            grad += np.einsum("pe,ped->pd", acc[:, 217:].astype(np.float64), demb)
            acc = acc.copy()
            acc[:, 217:] = 0.0
        if l == 0:
            grad += np.einsum("pe,ped->pd", acc[:, :39].astype(np.float64), demb)
            break
        gA = (64.0 * acc[:, : Ws[l - 1].shape[0]] * sigmoid_t(stash[l - 1])).astype(np.float32)
        if Ws[l - 1].shape[0] < 256:                                         # layer 3 has 217 outputs; its operand is 256 wide
            gA = np.concatenate([gA, np.zeros((gA.shape[0], 256 - gA.shape[1]), np.float32)], 1)[:, : Ws[l - 1].shape[0]]
    rel = lambda a_, b_: np.abs(a_ - b_).max() / max(1.0, np.abs(b_).max())
    assert rel(sdf, sdf_ref) < 3e-6 and rel(feat, feat_ref) < 3e-6, (rel(sdf, sdf_ref), rel(feat, feat_ref))
    assert rel(grad, grad_ref) < 1e-5, rel(grad, grad_ref)
