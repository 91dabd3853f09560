"""What is the reference's OWN numerical noise floor?  Run the oracle (== reference, oracle/ref_harness.py check)
twice: once as is, once with torch.exp perturbed by +-1 ulp on a random 30 % of the elements (a different but
equally valid libm, e.g. CUDA's expf vs SLEEF).  Prints how far z_vals and rendered pixels move."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hold_b200 import synth
from oracle import hold_oracle as O

torch.set_num_threads(8)
orig_exp = torch.exp
gen = torch.Generator().manual_seed(0)

def noisy_exp(x):
    y = orig_exp(x)
    r = torch.rand(y.shape, generator=gen)
    up = torch.nextafter(y, torch.full_like(y, float("inf")))
    dn = torch.nextafter(y, torch.full_like(y, float("-inf")))
    return torch.where(r < 0.15, up, torch.where(r > 0.85, dn, y))

for S, beta, add_tiny in [(32, 0.1, 1e-6), (128, 0.03, 1e-6), (128, 0.03, 1e-3)]:
    res = []
    for noisy in (False, True):
        torch.exp = noisy_exp if noisy else orig_exp
        sc = synth.make_scene(H=12, W=12, S=S, nodes=("right", "object"), seed=3)
        sc.sampler["add_tiny"] = add_tiny
        for n in sc.node_ids:
            sc.beta[n] = torch.tensor(beta)
        outs, _ = O.render_scene(sc)
        res.append(outs[0])
    torch.exp = orig_exp
    a, b = res
    for k, nid in enumerate(("right", "object")):
        dz = (a["nodes"][k]["z_vals"] - b["nodes"][k]["z_vals"]).abs()
        print(f"S={S} beta={beta} add_tiny={add_tiny} {nid}: z_vals frac>8e-4 {(dz > 8e-4).float().mean():.4f} max {dz.max():.3e}", end=" | ")
        for key in ("fg_rgb", "depth"):
            d = (a["render"][k][key] - b["render"][k][key]).abs()
            print(f"{key} frac>1e-4 {(d > 1e-4).float().mean():.3f} max {d.max():.2e}", end=" ")
        print()
    d = (a["render"]["comp"]["fg_rgb"] - b["render"]["comp"]["fg_rgb"]).abs()
    print(f"   comp fg_rgb frac>1e-4 {(d > 1e-4).float().mean():.3f} max {d.max():.2e}")
