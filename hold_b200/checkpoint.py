"""Loading a reference checkpoint into the mirror (drop-in boundary).

The reference saves `HOLD.state_dict()` (Lightning: ckpt["state_dict"]) with keys `model.nodes.<id>.<module>...` and
`model.background...` (hold/hold.py:30-60, hold/hold_net.py:20-51).  The mirror modules use the SAME names for everything
that carries learnable state on the hot path, so loading is a prefix strip plus a strict check on OUR keys:

    model.nodes.<id>.implicit_network.lin<k>.{weight_g,weight_v,bias}      -> Node.implicit_network
    model.nodes.<id>.rendering_network.lin<k>.*, .lin_pose.{weight,bias}   -> Node.rendering_network
    model.nodes.<id>.density.beta                                          -> Node.density
    model.nodes.<id>.params.<name>.weight   (model/generic/params.py)      -> Node.params (GenericParams mirror)
    model.nodes.object.frame_latent_encoder.weight                         -> Node.frame_latent_encoder
    model.background.{bg_implicit_network,bg_rendering_network}.lin<k>.*, .frame_latent_encoder.weight -> Background

    model.nodes.object.server.object_model.{obj_scale,norm_mat,v3d_cano}   -> ObjectServer.set_object_model (denorm_mat is
                                                                               rebuilt from norm_mat, object_model.py:27)

Ignored on purpose: the MANO server / deformer tensors (given to the mirror's constructors), the BARF embedder counters
(`embedder_obj.alpha_*`: eval() uses all-ones weights, render.py:43-47)."""
from __future__ import annotations

import torch
import torch.nn as nn

HAND_PARAMS = {"global_orient": 3, "pose": 45, "transl": 3, "betas": 10}     # model/mano/params.py
OBJECT_PARAMS = {"global_orient": 3, "transl": 3}                               # model/obj/params.py
IGNORED_SUBSTRINGS = (".server.", ".deformer.", ".object_model.", "embedder_obj.", ".embed_fn.", "alpha_iter", "alpha_max_iter")


class GenericParams(nn.Module):
    """model/generic/params.py:6-44: one nn.Embedding per pose parameter (betas: a single row), same attribute names."""

    def __init__(self, num_frames: int, params_dim: dict, node_id: str):
        super().__init__()
        self.num_frames, self.params_dim, self.node_id = num_frames, dict(params_dim), node_id
        for name, dim in params_dim.items():
            emb = nn.Embedding(1 if name == "betas" else num_frames, dim)
            emb.weight.data.fill_(0)
            emb.weight.requires_grad = False
            setattr(self, name, emb)

    def forward(self, frame_ids):
        out = {}
        for name in self.params_dim:
            ids = torch.zeros_like(frame_ids) if name == "betas" else frame_ids
            out[f"{self.node_id}.{name}"] = getattr(self, name)(ids)
        return out


OBJECT_MODEL_KEYS = ("obj_scale", "norm_mat", "v3d_cano")


def object_model_buffers(sd: dict, prefix: str = "model.") -> dict:
    """{node_id: {obj_scale, norm_mat, v3d_cano}} found under `<prefix>nodes.<id>.server.object_model.*`."""
    out = {}
    for k, v in sd.items():
        if k.startswith(prefix + "nodes.") and ".server.object_model." in k:
            nid = k[len(prefix) + len("nodes."):].split(".", 1)[0]
            name = k.rsplit(".", 1)[1]
            if name in OBJECT_MODEL_KEYS:
                out.setdefault(nid, {})[name] = v
    return out


def split_reference_state_dict(sd: dict, prefix: str = "model."):
    """-> ({node_id: {sub_key: tensor}}, {background sub_key: tensor}, [ignored keys]) from a reference state_dict."""
    nodes, bg, ignored = {}, {}, []
    for k, v in sd.items():
        if not k.startswith(prefix):
            ignored.append(k)
            continue
        r = k[len(prefix):]
        if any(s in "." + r for s in IGNORED_SUBSTRINGS):
            ignored.append(k)
        elif r.startswith("nodes."):
            _, nid, sub = r.split(".", 2)
            nodes.setdefault(nid, {})[sub] = v
        elif r.startswith("background."):
            bg[r[len("background."):]] = v
        else:
            ignored.append(k)
    return nodes, bg, ignored


def load_reference_state_dict(net, sd: dict, prefix: str = "model.", strict: bool = True):
    """Load a reference `state_dict` into a hold_b200 HOLDNet-like module (`net.nodes` ModuleDict, optional `net.background`)
    and push the weights to the device kernels (`sync_weights`).  strict: every parameter/buffer of OUR modules must be present
    in the checkpoint with the same shape (the reverse is not required: see the ignored list)."""
    nodes, bg, ignored = split_reference_state_dict(sd, prefix)
    missing, loaded = [], 0
    for nid, node in net.nodes.items():
        own = node.state_dict()
        src = nodes.get(nid, {})
        for k, t in own.items():
            if k in src and tuple(src[k].shape) == tuple(t.shape):
                t.copy_(src[k].to(t.device, t.dtype))
                loaded += 1
            else:
                missing.append(f"nodes.{nid}.{k}")
    # ObjectModel buffers: not parameters of the mirror, but they decide the object's transform (obj_tfs, verts)
    for nid, bufs in object_model_buffers(sd, prefix).items():
        node = net.nodes[nid] if nid in net.nodes else None
        if node is None or not hasattr(getattr(node, "server", None), "set_object_model"):
            continue
        if "v3d_cano" in bufs and tuple(bufs["v3d_cano"].shape) != tuple(node.server.v3d_cano.shape) and strict:
            raise KeyError(f"nodes.{nid}.server.object_model.v3d_cano is {tuple(bufs['v3d_cano'].shape)}, the mirror holds "
                           f"{tuple(node.server.v3d_cano.shape)}")
        node.server.set_object_model(obj_scale=bufs.get("obj_scale"), norm_mat=bufs.get("norm_mat"), v3d_cano=bufs.get("v3d_cano"))
        loaded += len(bufs)
    if getattr(net, "background", None) is not None:
        own = net.background.state_dict()
        for k, t in own.items():
            if k in bg and tuple(bg[k].shape) == tuple(t.shape):
                t.copy_(bg[k].to(t.device, t.dtype))
                loaded += 1
            else:
                missing.append(f"background.{k}")
    if strict and missing:
        raise KeyError(f"reference checkpoint lacks {len(missing)} tensors the mirror needs, e.g. {missing[:5]}")
    if hasattr(net, "sync_weights"):
        net.sync_weights()
    return {"loaded": loaded, "missing": missing, "ignored": ignored}
