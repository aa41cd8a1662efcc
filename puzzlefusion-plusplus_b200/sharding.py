"""Object-level data parallelism: one process per GPU, no data-path collective.

Objects are independent (no cross-object op anywhere in test_step), so the only exchange is one
all_gather of the per-object [B_local, 4] metric block at the end -- the B200-native replacement of
Lightning's ``self.log(..., sync_dist=True)`` reductions (auto_aggl.py:366-369).  Load balance matters
more than the collective: objects leave the outer loop after different iteration counts, so they are
dealt to ranks in serpentine order of decreasing fragment count.
"""
import torch
import torch.distributed as dist


def shard_objects(num_parts, rank, world):
    """Indices of the objects this rank owns."""
    order = sorted(range(len(num_parts)), key=lambda i: (-int(num_parts[i]), i))
    mine = []
    for k, i in enumerate(order):  # serpentine deal: 0..w-1, w-1..0, ...
        r, c = divmod(k, world)
        if (c if r % 2 == 0 else world - 1 - c) == rank:
            mine.append(i)
    return sorted(mine)


def gather_metrics(block, mine, total):
    """all_gather the [len(mine), 4] blocks (padded to the largest shard) -> [total, 4] in dataset order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = torch.zeros(total, block.shape[1], dtype=block.dtype, device=block.device)
        out[torch.as_tensor(mine, dtype=torch.long)] = block
        return out
    world = dist.get_world_size()
    cap = (total + world - 1) // world
    pad = torch.full((cap, block.shape[1] + 1), -1.0, dtype=block.dtype, device=block.device)
    pad[:len(mine), 0] = torch.as_tensor(mine, dtype=block.dtype, device=block.device)
    pad[:len(mine), 1:] = block
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    allp = torch.cat(gathered)
    allp = allp[allp[:, 0] >= 0]
    out = torch.zeros(total, block.shape[1], dtype=block.dtype, device=block.device)
    out[allp[:, 0].long()] = allp[:, 1:]
    return out
