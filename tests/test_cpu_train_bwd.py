"""hold_b200/train_algo.py (the algebra of the training backward: SDF net with its input gradient as an output, incl. the
second-order terms; colour net) against torch.autograd over the oracle, in float64 on the CPU — hand and object nets, reference
initialisation and perturbed weights, with BARF weights on the object's embedding."""
import math

import pytest
import torch


def _folded(sd, n, hand):
    from oracle import hold_oracle as O

    W = [O.wn(sd, f"lin{l}").clone() for l in range(n)]
    b = [sd[f"lin{l}.bias"].clone() for l in range(n)]
    if hand and n == 9:
        W[0] = W[0][:, :39].clone()   # the 45 pose-condition columns are multiplied by zero (shape_net.py:104-106)
    return W, b


@pytest.mark.parametrize("kind,perturb,barf", [("hand", 0.0, False), ("object", 0.02, True), ("hand", 0.05, False)])
def test_sdf_backward_matches_autograd(kind, perturb, barf):
    from hold_b200 import synth, train_algo as T
    from oracle import hold_oracle as O

    torch.manual_seed(1)
    sd = {k: v.double() for k, v in synth.make_sdf_state(kind, 3, 0.6, perturb).items()}
    P = 300
    x = (torch.rand(P, 3, dtype=torch.float64) - 0.5) * 1.6
    ew = O.barf_weights(2.6).double() if barf else None
    W, b = _folded(sd, 9, kind == "hand")
    Wp = [w.clone().requires_grad_(True) for w in W]
    bp = [v.clone().requires_grad_(True) for v in b]
    xg = x.clone().requires_grad_(True)

    # autograd reference: the oracle's forward with explicit folded weights, gradient with create_graph (volsdf_utils.py:89-96,129)
    def fwd(xx):
        e = O.embed(xx, 6, ew)
        h = e
        for l in range(9):
            if l == 4:
                h = torch.cat([h, e], 1) / math.sqrt(2)
            h = torch.nn.functional.linear(h, Wp[l], bp[l])
            if l < 8:
                h = torch.nn.functional.softplus(h, beta=100)
        return h

    out = fwd(xg)
    g_ref = torch.autograd.grad(out[:, 0].sum(), xg, create_graph=True)[0]
    # check the oracle equivalence of this explicit forward once
    cond = torch.zeros(P, 45, dtype=torch.float64) if kind == "hand" else None
    assert (O.sdf_mlp(x, sd, cond, ew) - out.detach()).abs().max() < 1e-12
    d_sdf, d_feat, d_g = torch.randn(P, dtype=torch.float64), torch.randn(P, 256, dtype=torch.float64) * 0.1, torch.randn(P, 3, dtype=torch.float64)
    loss = (out[:, 0] * d_sdf).sum() + (out[:, 1:] * d_feat).sum() + (g_ref * d_g).sum()
    grads = torch.autograd.grad(loss, [xg] + Wp + bp)
    Wk = [w.clone() for w in W]
    Wk[4] = Wk[4] / math.sqrt(2)          # the algorithm's matrices carry the skip's 1/sqrt 2 (as the packed weight images do)
    ops = T.TorchOps(Wk, b)
    sdf, feat, g, st = T.sdf_forward(ops, x, ew)
    assert (sdf - out[:, 0].detach()).abs().max() < 1e-12 and (feat - out[:, 1:].detach()).abs().max() < 1e-12
    assert ((g - g_ref.detach()).abs().max() / g_ref.detach().abs().max()) < 1e-8   # nn.Softplus switches to the identity above 100 z = 20 (2e-9)
    d_x, dW, db = T.sdf_backward(ops, st, d_sdf, d_feat, d_g)
    dW[4] = dW[4] / math.sqrt(2)
    rel = lambda a, r: ((a - r).abs().max() / r.abs().max().clamp_min(1e-30)).item()
    assert rel(d_x, grads[0]) < 1e-7, rel(d_x, grads[0])
    for l in range(9):
        assert rel(dW[l], grads[1 + l]) < 1e-7, (l, rel(dW[l], grads[1 + l]))
        assert rel(db[l], grads[10 + l]) < 1e-7, (l, rel(db[l], grads[10 + l]))
    # first-order only (no gradient seed) must work too
    d_x1, dW1, _ = T.sdf_backward(ops, st, d_sdf, None, None)
    g1 = torch.autograd.grad((fwd(xg)[:, 0] * d_sdf).sum(), [xg, Wp[2]])
    assert rel(d_x1, g1[0]) < 1e-7 and rel(dW1[2], g1[1]) < 1e-7


@pytest.mark.parametrize("kind", ["hand", "object"])
def test_rgb_backward_matches_autograd(kind):
    from hold_b200 import synth, train_algo as T
    from oracle import hold_oracle as O

    torch.manual_seed(2)
    sd = {k: v.double() for k, v in synth.make_rgb_state(kind, 3).items()}
    P = 200
    K0 = 270 if kind == "hand" else 302
    inp = torch.randn(P, K0, dtype=torch.float64) * 0.5
    W, b = _folded(sd, 5, False)
    Wp = [w.clone().requires_grad_(True) for w in W]
    bp = [v.clone().requires_grad_(True) for v in b]
    ig = inp.clone().requires_grad_(True)
    h = ig
    for l in range(5):
        h = torch.nn.functional.linear(h, Wp[l], bp[l])
        if l < 4:
            h = torch.relu(h)
    ref = torch.sigmoid(h)
    d_rgb = torch.randn(P, 3, dtype=torch.float64)
    grads = torch.autograd.grad((ref * d_rgb).sum(), [ig] + Wp + bp)
    ops = T.TorchOps(W, b)
    rgb, st = T.rgb_forward(ops, inp)
    assert (rgb - ref.detach()).abs().max() < 1e-12
    d_in, dW, db = T.rgb_backward(ops, st, d_rgb)
    rel = lambda a, r: ((a - r).abs().max() / r.abs().max().clamp_min(1e-30)).item()
    assert rel(d_in, grads[0]) < 1e-7
    for l in range(5):
        assert rel(dW[l], grads[1 + l]) < 1e-7 and rel(db[l], grads[6 + l]) < 1e-7
