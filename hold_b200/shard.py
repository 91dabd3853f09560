"""Multi-GPU partitioning of the path: rays/frames are independent units, so ranks take disjoint
ray ranges with NO data-path collective (render.py has no gradients).  Shards are aligned to the
reference's 512-pixel render chunks (datasets/eval_datasets.py:13) so that a sharded render reproduces the
reference's own chunked semantics of the batch-global sampler flag (engine/ray_sampler.py:244)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int, align: int = 512) -> tuple[int, int]:
    """[start, stop) of `n` units for `rank` of `world`, boundaries on multiples of `align` (last shard takes the tail)."""
    assert 0 <= rank < world and n >= 0 and align >= 1
    blocks = (n + align - 1) // align
    per, rem = divmod(blocks, world)
    b0 = rank * per + min(rank, rem)
    b1 = b0 + per + (1 if rank < rem else 0)
    return min(b0 * align, n), min(b1 * align, n)


def shard_frames(n_frames: int, rank: int, world: int) -> list[int]:
    """Round-robin frames, the reference's own `--agent_id` farming idea (datasets/eval_datasets.py:43-50)."""
    return list(range(rank, n_frames, world))
