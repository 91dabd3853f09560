"""render.py-style batch render (BASELINE.json configs[3]): N frames of a scene, frames farmed round-robin over the ranks of
one box (the reference's own `--agent_id` idea, datasets/eval_datasets.py:43-50), each rank renders whole frames with the
full forward (foreground nodes + NeRF++ background + composite) and rank 0 assembles the image stack.  No data-path
collective; one gather for assembly.

  python tools/render_frames.py --frames 8 --size 128                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
      tools/render_frames.py --frames 200 --size 512 --out /tmp/frames.npy   # 8 GPUs

Synthetic scene (hold_b200.synth) unless --ckpt gives a reference checkpoint to load into the same modules
(hold_b200.checkpoint; MANO tensors are still the synthetic ones: the licensed MANO pickle is not redistributable)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--nodes", default="right,object")
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--mode", default="tc", choices=["tc", "fp32"])
    a = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from hold_b200 import capi, checkpoint, scene_io, shard, synth
    from hold_b200.model import HOLDNet

    dev = torch.device("cuda", local)
    ctx = capi.Context(local)
    sc = synth.make_scene(H=a.size, W=a.size, S=a.samples, nodes=tuple(a.nodes.split(",")), B=1, seed=0)
    net = scene_io.build_net(sc, ctx, capi.MLP_TC if a.mode == "tc" else capi.MLP_FP32)
    bg, _, _ = scene_io.build_background(sc, ctx)
    full = HOLDNet(ctx, dict(net.nodes), background=bg)
    if a.ckpt:
        info = checkpoint.load_reference_state_dict(full, torch.load(a.ckpt, map_location="cpu")["state_dict"], strict=False)
        if rank == 0:
            print(f"checkpoint: {info['loaded']} tensors loaded, {len(info['missing'])} missing, {len(info['ignored'])} ignored")
    mine = shard.shard_frames(a.frames, rank, world)
    inp = scene_io.scene_input(sc, dev)
    imgs = torch.zeros(len(mine), a.size * a.size, 3, device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    for k, f in enumerate(mine):
        # a different camera per frame: rotate the synthetic camera about the scene's up axis
        ang = 2.0 * np.pi * f / max(a.frames, 1)
        Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0, 0], [np.sin(ang), np.cos(ang), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32, device=dev)
        fi = dict(inp)
        fi["extrinsics"] = Rz[None] @ inp["extrinsics"]
        out = full(fi)
        imgs[k] = out["rgb"]
    ctx.check()
    torch.cuda.synchronize()
    dt = time.time() - t0
    rays = len(mine) * a.size * a.size
    line = f"rank {rank}: {len(mine)} frames, {rays / max(dt, 1e-9):.0f} rays/s"
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sizes = [len(shard.shard_frames(a.frames, r, world)) for r in range(world)]
        pad = torch.zeros(max(sizes), a.size * a.size, 3, device=dev)
        pad[: len(mine)] = imgs
        allf = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, allf, dst=0)
        if rank == 0:
            stack = torch.zeros(a.frames, a.size * a.size, 3, device=dev)
            for r in range(world):
                for k, f in enumerate(shard.shard_frames(a.frames, r, world)):
                    stack[f] = allf[r][k]
            imgs, dt = stack, t.item()
    print(line, flush=True)
    if rank == 0:
        print(f"total: {a.frames} frames of {a.size}x{a.size} in {dt:.2f} s = {a.frames * a.size * a.size / dt:.0f} rays/s over {world} GPU(s)")
        if a.out:
            np.save(a.out, imgs.reshape(-1, a.size, a.size, 3).cpu().numpy())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
