"""CPU-only host logic: synthetic scenes are deterministic, the mirror modules keep the reference's state_dict key
format, the FLOP accounting matches SURVEY §8d, and ray sharding (world_size 2, gloo) partitions the rays exactly."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic():
    from hold_b200 import synth

    a, b = synth.make_scene(H=8, W=8, S=32, seed=5), synth.make_scene(H=8, W=8, S=32, seed=5)
    assert torch.equal(a.uv, b.uv) and torch.equal(a.mano["right"]["v_template"], b.mano["right"]["v_template"])
    for k in a.sdf_state["right"]:
        assert torch.equal(a.sdf_state["right"][k], b.sdf_state["right"][k])
    m = a.mano["right"]
    assert m["v_template"].shape == (778, 3) and m["shapedirs"].shape == (778, 3, 10) and m["posedirs"].shape == (135, 2334)
    assert torch.allclose(m["lbs_weights"].sum(1), torch.ones(778), atol=1e-6)
    assert torch.allclose(m["J_regressor"].sum(1), torch.ones(16), atol=1e-5)
    assert a.sampler == dict(near=0.0, N_samples=16, N_samples_eval=32, N_samples_extra=8, eps=0.1, beta_iters=10,
                             max_total_iters=5, add_tiny=1e-6)


def test_mirror_state_dict_keys_match_reference_format():
    from hold_b200 import synth
    from hold_b200.model import ImplicitNet, LaplaceDensity, RenderingNet

    for kind in ("hand", "object"):
        net, rgb = ImplicitNet(kind), RenderingNet(kind)
        keys = set(net.state_dict())
        assert keys == {f"lin{l}.{p}" for l in range(9) for p in ("weight_g", "weight_v", "bias")}
        net.load_state_dict(synth.make_sdf_state(kind, 0, 0.5), strict=True)
        rgb.load_state_dict(synth.make_rgb_state(kind, 0), strict=True)
        assert net.lin3.weight_v.shape == (217, 256) and net.lin8.weight_v.shape == (257, 256)
        assert rgb.lin0.weight_v.shape[1] == (270 if kind == "hand" else 302)
    assert list(LaplaceDensity().state_dict()) == ["beta"]


def test_flop_accounting_matches_survey():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.MAC_SDF_FULL == 524_544
    assert abs(bench.flops_per_ray(5, ("right",)) / 1e6 - 916.0) < 1.0        # 916 MFLOP per hand-ray
    assert abs(bench.flops_per_ray(5, ("right", "object")) / 1e9 - 1.834) < 0.002


def test_shard_range_partitions_exactly():
    from hold_b200.shard import shard_frames, shard_range

    for n in (0, 1, 511, 512, 513, 4096, 262144, 262145):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(s % 512 == 0 for s, _ in spans if s < n)
    assert sorted(sum((shard_frames(10, r, 3) for r in range(3)), [])) == list(range(10))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    from hold_b200.shard import shard_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4096 + 37
    s, e = shard_range(n, rank, world)
    mine = torch.zeros(n)
    mine[s:e] = 1.0
    dist.all_reduce(mine)                       # every ray owned by exactly one rank
    t = torch.tensor([float(e - s)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)    # the bench's max-over-ranks reduction
    # one flat-bucket all-reduce over several gradient tensors (frame-sharded pose refinement)
    from hold_b200.shard import allreduce_flat_, gather_rays
    g1, g2 = torch.full((10,), float(rank + 1)), torch.full((2, 3), 10.0 * (rank + 1))
    allreduce_flat_([g1, None, g2])
    ok_ar = bool((g1 == 3.0).all() and (g2 == 30.0).all())
    # gradient buckets keep the same layout on every rank even when a rank has no gradient for a parameter (a node none of its
    # rays hit): rank 0 lacks p2's gradient, rank 1 lacks p1's
    from hold_b200.shard import allreduce_grads_
    p1, p2, p3 = (torch.nn.Parameter(torch.zeros(5)) for _ in range(3))
    if rank == 0:
        p1.grad = torch.full((5,), 2.0)
    else:
        p2.grad = torch.full((5,), 7.0)
    p3.grad = torch.full((5,), float(rank + 1))
    allreduce_grads_([p1, p2, p3], average=True)
    ok_ar = ok_ar and bool((p1.grad == 1.0).all() and (p2.grad == 3.5).all() and (p3.grad == 1.5).all())
    # image assembly on rank 0
    local = torch.arange(s, e, dtype=torch.float32)[:, None].repeat(1, 3)
    full = gather_rays(local, n, rank, world)
    ok_g = (full is None) if rank != 0 else bool(torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32)))
    q.put((rank, bool((mine == 1).all()) and ok_ar and ok_g, t.item()))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok, _ in res) and all(p.exitcode == 0 for p in ps)
    assert {r for r, _, _ in res} == {0, 1}


def test_reference_checkpoint_key_mapping():
    """hold_b200.checkpoint: a state_dict laid out like the reference's (hold/hold.py + hold_net.py naming) loads into modules
    with the mirror's names; server / deformer buffers are ignored; strictness is on OUR keys."""
    import torch.nn as nn

    from hold_b200 import checkpoint as ck
    from hold_b200 import synth
    from hold_b200.model import ImplicitNet, LaplaceDensity, RenderingNet

    class FakeNode(nn.Module):           # the mirror Node's learnable sub-modules, without a device context
        def __init__(self, kind, n_frames):
            super().__init__()
            self.implicit_network, self.rendering_network, self.density = ImplicitNet(kind), RenderingNet(kind), LaplaceDensity()
            self.params = ck.GenericParams(n_frames, ck.HAND_PARAMS if kind == "hand" else ck.OBJECT_PARAMS, "right" if kind == "hand" else "object")
            if kind == "object":
                self.frame_latent_encoder = nn.Embedding(n_frames, 32)

    class FakeNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.nodes = nn.ModuleDict({"right": FakeNode("hand", 7), "object": FakeNode("object", 7)})
            self.synced = 0

        def sync_weights(self):
            self.synced += 1

    # a "reference" state dict: same tensors under model.*, plus buffers the mirror must ignore
    src = FakeNet()
    for nid, kind in (("right", "hand"), ("object", "object")):
        src.nodes[nid].implicit_network.load_state_dict(synth.make_sdf_state(kind, 3, 0.5))
        src.nodes[nid].rendering_network.load_state_dict(synth.make_rgb_state(kind, 3))
        src.nodes[nid].params.global_orient.weight.data.normal_()
    sd = {"model." + k: v.clone() for k, v in src.state_dict().items()}
    sd["model.nodes.right.server.verts_c"] = torch.zeros(1, 778, 3)
    sd["model.nodes.object.server.object_model.v3d_cano"] = torch.zeros(10, 3)
    sd["model.nodes.object.implicit_network.embedder_obj.alpha_iter"] = torch.tensor(5)
    sd["loss.some_buffer"] = torch.zeros(1)
    dst = FakeNet()
    info = ck.load_reference_state_dict(dst, sd)
    assert dst.synced == 1 and not info["missing"] and len(info["ignored"]) == 4
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    p = dst.nodes["right"].params(torch.tensor([0, 3]))
    assert set(p) == {"right.global_orient", "right.pose", "right.transl", "right.betas"} and p["right.betas"].shape == (2, 10)
    # ObjectModel buffers (model/obj/object_model.py:23-27) reach the object server's set_object_model
    class FakeServer:
        def __init__(self):
            self.v3d_cano, self.got = torch.zeros(10, 3), None

        def set_object_model(self, obj_scale=None, norm_mat=None, v3d_cano=None):
            self.got = dict(obj_scale=obj_scale, norm_mat=norm_mat, v3d_cano=v3d_cano)

    dst2 = FakeNet()
    object.__setattr__(dst2.nodes["object"], "server", FakeServer())
    sd["model.nodes.object.server.object_model.obj_scale"] = torch.tensor([0.83])
    sd["model.nodes.object.server.object_model.norm_mat"] = torch.eye(4) * 2.0
    sd["model.nodes.object.server.object_model.denorm_mat"] = torch.eye(4) * 0.5
    ck.load_reference_state_dict(dst2, sd)
    got = dst2.nodes["object"].server.got
    assert got is not None and float(got["obj_scale"]) == pytest.approx(0.83) and got["norm_mat"][0, 0] == 2.0 and got["v3d_cano"].shape == (10, 3)
    del sd["model.nodes.right.density.beta"]
    with pytest.raises(KeyError):
        ck.load_reference_state_dict(FakeNet(), sd)

