"""Round-2 kernel variants (all behind environment switches, none had a hardware run when this was written): each is run in a
fresh interpreter with its switch set and held to the same bar as the default tensor-core kernels (tests/test_gpu_tc.py:
sdf / feature / gradient within 1e-4 of the exact-fp32 CUDA-core kernels).  Non-strict xfail: XPASS = ready for A/B timing
(tools/exp_matrix.py)."""
import os

import pytest
import torch

# Opt-in (HOLD_RUN_VARIANTS=1, set by tools/round2_first_call.sh): these drive tcgen05 / mbarrier protocols that never ran on
# hardware; they execute in separate processes with bounded waits, but a first run belongs in a supervised GPU call, not in an
# unattended suite that is followed by the measurements.
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("HOLD_RUN_VARIANTS") != "1", reason="opt-in: HOLD_RUN_VARIANTS=1"),
              pytest.mark.xfail(strict=False, reason="first hardware run pending (written after the round's GPU budget was spent)")]


def impl_sdf_kernels_match_fp32(ctx):
    from hold_b200 import capi, scene_io, synth

    dev = torch.device("cuda", 0)
    sc = synth.make_scene(H=8, W=8, S=32, nodes=("right", "object"), seed=4)
    P = 5000   # not a multiple of the tile sizes
    x = ((torch.rand(P, 3, generator=torch.Generator().manual_seed(9)) - 0.5) * 1.6).to(dev)
    res = {}
    for mode in (capi.MLP_FP32, capi.MLP_TC):
        net = scene_io.build_net(sc, ctx, mode)
        for nid in sc.node_ids:
            node = net.nodes[nid]
            s0 = torch.empty(P, device=dev)
            capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(x), None, capi.ptr(s0), None, None, capi.stream_ptr()))
            s1, g1, f1 = torch.empty(P, device=dev), torch.empty(P, 3, device=dev), torch.empty(P, 256, device=dev)
            capi.check(capi.lib().hold_sdf_eval(ctx.h, node.slot, P, capi.ptr(x), None, capi.ptr(s1), capi.ptr(g1), capi.ptr(f1), capi.stream_ptr()))
            ctx.check()
            res[(mode, nid)] = [t.cpu() for t in (s0, s1, g1, f1)]
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1.0)).item()
    for nid in sc.node_ids:
        for name, a, b in zip(("sdf_only", "sdf", "grad", "feat"), res[(capi.MLP_TC, nid)], res[(capi.MLP_FP32, nid)]):
            e = rel(a, b)
            print(f"{nid}.{name}: {e:.2e}")
            assert e < 1e-4, f"{nid}.{name}: {e:.2e}"


def impl_render_matches_default(ctx):
    """whole foreground frame with the switch set vs the oracle at the e2e bar of tests/test_gpu_e2e.py (tensor-core mode)"""
    from hold_b200 import capi, scene_io, synth
    from oracle import hold_oracle as O

    sc = synth.make_scene(H=8, W=8, S=32, nodes=("right", "object"), seed=4)
    net = scene_io.build_net(sc, ctx, capi.MLP_TC)
    out = net.forward_fg(scene_io.scene_input(sc, torch.device("cuda", 0)))
    ctx.check()
    ref, _ = O.render_scene(sc, stable_ties=True)
    for k, nid in enumerate(sc.node_ids):
        for key in ("fg_rgb", "depth", "normal", "mask_prob"):
            a, b = out[f"{nid}.{key}"].cpu().reshape(ref[0]["render"][k][key].shape), ref[0]["render"][k][key]
            d = (a - b).abs()
            frac = (d <= 1e-4 * max(1.0, b.abs().max().item())).float().mean().item()
            assert frac >= 0.95 and d.max().item() < 1e-2, f"{nid}.{key}: within {frac:.3f}, max {d.max().item():.2e}"


@pytest.mark.parametrize("name,env", [
    ("lean", dict(HOLD_TC_LEAN="1")),
    ("lean_tstash", dict(HOLD_TC_LEAN="1", HOLD_TC_DBG="64")),
    ("fast", dict(HOLD_TC_FAST="1")),
    ("pair_coarse", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="32")),
    ("pair_light", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="160")),
    ("pair_wide", dict(HOLD_TC_PAIR="1", HOLD_TC_DBG="384")),
    ("fast_replicas", dict(HOLD_TC_FAST="1", HOLD_TC_WCOPIES="4")),
])
def test_sdf_kernel_variant(isolated, name, env):
    isolated("tests/test_gpu_variants.py", "impl_sdf_kernels_match_fp32", env=env)


@pytest.mark.parametrize("name,env", [
    ("knn_filter", dict(HOLD_KNN_FILTER="1")),
    ("knn_occ", dict(HOLD_KNN_OCC="1")),
    ("fast", dict(HOLD_TC_FAST="1")),
])
def test_render_variant(isolated, name, env):
    isolated("tests/test_gpu_variants.py", "impl_render_matches_default", env=env)
