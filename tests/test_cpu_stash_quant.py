"""The reverse-mode SDF kernel (k_mlp_tc<MLP_SDF_REV>, hold_b200/csrc/mlp_tc.cuh) stashes softplus'(z_l) of its 8 forward
layers as unorm16 (512 KB per CTA instead of 1 MB: the stash of the whole grid stays L2-resident).  Two CPU checks of that
choice: (1) the bit tricks of the encode / decode (magic-number add, byte permute into the mantissa of 2^23) restated in numpy
are an exact round-to-nearest unorm16; (2) the gradient d sdf / d x_c computed by the manual reverse pass with quantised
softplus' stays within 2e-5 of its scale for the reference-initialised and a perturbed net (fp16 storage would give 2e-4)."""
import math

import numpy as np
import torch


def _pack2(s0, s1):
    # fmaf(s, 65535, 2^23): one rounding (the float64 product and sum are exact: 24 + 16 bits)
    b0 = (np.float64(s0) * 65535.0 + 8388608.0).astype(np.float32).view(np.uint32)
    b1 = (np.float64(s1) * 65535.0 + 8388608.0).astype(np.float32).view(np.uint32)
    return (b0 & 0xFFFF) | ((b1 & 0xFFFF) << 16)     # __byte_perm(b0, b1, 0x5410)


def _lo(w):
    return ((w & 0xFFFF) | np.uint32(0x4B000000)).view(np.float32) - np.float32(8388608.0)   # __byte_perm(w, 0x4B000000, 0x7410)


def _hi(w):
    return ((w >> 16) | np.uint32(0x4B000000)).view(np.float32) - np.float32(8388608.0)      # __byte_perm(w, 0x4B000000, 0x7432)


def test_unorm16_bit_tricks():
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.random(100000), [0.0, 1.0, 0.5, 1e-7, 1 - 1e-7, 0.5 / 65535, 1.5 / 65535]]).astype(np.float32)
    if s.size % 2:
        s = s[:-1]
    w = _pack2(s[0::2], s[1::2])
    q = np.empty_like(s)
    q[0::2], q[1::2] = _lo(w), _hi(w)
    assert q.min() >= 0 and q.max() <= 65535 and np.all(q == np.round(q))
    # fmaf(s, 65535, 2^23) rounds once, in fp32, to an integer: the nearest integer of the exact product
    exact = s.astype(np.float64) * 65535.0
    assert np.all(np.abs(q - exact) <= 0.5 + 1e-9)
    assert np.abs(q / 65535.0 - s).max() <= 0.5 / 65535 + 1e-9


def test_gradient_error_of_quantised_stash():
    from hold_b200 import synth
    from oracle import hold_oracle as O

    torch.manual_seed(0)
    for kind, perturb in (("hand", 0.0), ("object", 0.0), ("hand", 0.02)):
        sd = {k: v.double() for k, v in synth.make_sdf_state(kind, 3, 0.6, perturb).items()}
        x = (torch.rand(1500, 3, dtype=torch.float64) - 0.5) * 1.6
        W = [O.wn(sd, f"lin{l}") for l in range(9)]
        b = [sd[f"lin{l}.bias"] for l in range(9)]
        if kind == "hand":
            W[0] = W[0][:, :39]   # the pose condition is multiplied by zero (shape_net.py:104-106)
        e = O.embed(x)

        def forward(q):
            a, sig = e, []
            for l in range(8):
                if l == 4:
                    a = torch.cat([a, e], 1) / math.sqrt(2)
                z = a @ W[l].T + b[l]
                s = torch.sigmoid(100 * z)
                sig.append(torch.round(s * q) / q if q else s)
                a = torch.nn.functional.softplus(z, beta=100)
            return sig

        def embed_chain(ge):
            xg = x.clone().requires_grad_(True)
            return torch.autograd.grad((O.embed(xg) * ge).sum(), xg)[0]

        def backward(sig):
            g = W[8][0][None, :] * sig[7]
            gx = torch.zeros_like(x)
            for l in range(7, 0, -1):
                ga = g @ W[l]
                if l == 4:
                    ga = ga / math.sqrt(2)
                    gx = gx + embed_chain(ga[:, 217:])
                    ga = ga[:, :217]
                g = ga * sig[l - 1]
            return gx + embed_chain(g @ W[0])

        g0 = backward(forward(None))
        xg = x.clone().requires_grad_(True)
        out = O.sdf_mlp(xg, sd, torch.zeros(x.shape[0], 45, dtype=torch.float64) if kind == "hand" else None)
        ga = torch.autograd.grad(out[:, 0].sum(), xg)[0]
        assert (g0 - ga).abs().max().item() < 1e-7, "manual reverse pass == autograd"
        gq = backward(forward(65535.0))
        rel = ((gq - g0).abs().max() / g0.abs().max()).item()
        print(f"{kind} perturb {perturb}: gradient error of the unorm16 stash {rel:.2e} of max|g|")
        assert rel <= 2e-5
