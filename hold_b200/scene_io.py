"""Glue between hold_b200.synth.SynthScene and the model mirror: builds a HOLDNet with the scene's
weights/rig and the reference-style input dict (`hold.py:116-121`, `mano_node.py:72-79`)."""
from __future__ import annotations

import torch

from . import capi
from .model import HOLDNet, Node


def build_background(sc, ctx, seed=0, mlp_mode=capi.MLP_FP32):
    """Background with hold_b200.synth.make_bg_state weights and a seeded frame-code table."""
    from . import synth
    from .model import Background

    dev = torch.device("cuda", ctx.device)
    bg = Background(ctx, sc.B, mlp_mode).to(dev)
    sdf_sd, rgb_sd = synth.make_bg_state(seed)
    bg.bg_implicit_network.load_state_dict(sdf_sd, strict=True)
    bg.bg_rendering_network.load_state_dict(rgb_sd, strict=True)
    g = torch.Generator().manual_seed(5000 + seed)
    bg.frame_latent_encoder.weight.data.copy_((0.1 * torch.randn(sc.B, 32, generator=g)).to(dev))
    bg.sync_weights()
    return bg, sdf_sd, rgb_sd


def build_net(sc, ctx, mlp_mode=capi.MLP_FP32):
    dev = torch.device("cuda", ctx.device)
    nodes = {}
    for slot, nid in enumerate(sc.node_ids):
        if nid in ("right", "left"):
            node = Node(ctx, slot, nid, sc.sampler, sc.bounding_sphere, n_frames=sc.B, mano=sc.mano[nid], betas=sc.betas[nid],
                        mlp_mode=mlp_mode, beta=float(sc.beta[nid]))
        else:
            node = Node(ctx, slot, nid, sc.sampler, sc.bounding_sphere, n_frames=sc.B, obj_pts=sc.obj_pts_cano, mlp_mode=mlp_mode,
                        beta=float(sc.beta[nid]))
        node.to(dev)
        node.implicit_network.load_state_dict(sc.sdf_state[nid], strict=True)
        rsd = dict(sc.rgb_state[nid])
        node.rendering_network.load_state_dict(rsd, strict=True)
        if nid == "object":
            node.frame_latent_encoder.weight.data.copy_(sc.time_code.to(dev))
        node.sync_weights()
        nodes[nid] = node
    return HOLDNet(ctx, nodes)


def scene_input(sc, dev, ray_ids=None):
    """Reference-style batch dict.  ray_ids: flat indices into one frame's H*W pixels (B must be 1) or None."""
    uv = sc.uv
    if ray_ids is not None:
        assert sc.B == 1
        uv = uv[:, ray_ids]
    inp = {"uv": uv.to(dev), "intrinsics": sc.intrinsics.to(dev), "extrinsics": sc.extrinsics.to(dev), "idx": sc.frame_idx.to(dev)}
    B = sc.B
    for nid in sc.node_ids:
        p = sc.params[nid]
        inp[f"{nid}.params"] = torch.full((B, 1), float(sc.scene_scale), device=dev)
        inp[f"{nid}.global_orient"] = p["global_orient"].to(dev)
        inp[f"{nid}.transl"] = p["transl"].to(dev)
        if nid != "object":
            inp[f"{nid}.pose"] = p["pose"].to(dev)
            inp[f"{nid}.betas"] = sc.betas[nid][None].repeat(B, 1).to(dev)
            inp[f"{nid}.full_pose"] = torch.cat([p["global_orient"], p["pose"]], 1).to(dev)
    return inp
