"""Frame-range sharding across GPUs.

Frames are independent (own first sample, own CRCs, even length), so a recording or a batch of files is split
at frame boundaries, every rank encodes / decodes its own range with no data-path collective, and the only
exchange is one integer per rank -- the shard's compressed size -- from which every rank derives its base
offset in the concatenated stream.  torch.distributed is plumbing: NCCL over NVLink on GPUs, gloo in CPU tests.
"""
from typing import List, Sequence, Tuple


def shard_frames(n_samples: int, samples_per_frame: int, rank: int, world: int) -> Tuple[int, int]:
    """Sample range [s0, s1) of rank's shard: frames [floor(rank*F/world), floor((rank+1)*F/world))."""
    frames = (n_samples + samples_per_frame - 1) // samples_per_frame
    f0 = frames * rank // world
    f1 = frames * (rank + 1) // world
    return min(n_samples, f0 * samples_per_frame), min(n_samples, f1 * samples_per_frame)


def deal_files(frames_per_file: Sequence[int], world: int) -> List[List[int]]:
    """Whole files dealt by cumulative frame count: rank r gets the files whose first frame falls in its
    equal share of the total.  Every file keeps its own archive prefix, so files are never split."""
    total = sum(frames_per_file)
    out: List[List[int]] = [[] for _ in range(world)]
    acc = 0
    for i, nf in enumerate(frames_per_file):
        r = min(world - 1, acc * world // total) if total else 0
        out[r].append(i)
        acc += nf
    return out


def exchange_sizes(local_size: int, dist=None, device=None) -> Tuple[List[int], int]:
    """All-gather one int64 per rank (the shard's compressed bytes); returns (sizes, this rank's base offset).
    With NCCL the tensor lives on the rank's GPU and the gather runs over NVLink; no other collective exists
    in the codec path."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(local_size)], 0
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(local_size)], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, mine)
    lst = [int(x) for x in sizes.tolist()]
    return lst, sum(lst[:rank])


def exchange_sizes_begin(local_size: int, dist=None, device=None):
    """Start the all-gather of exchange_sizes without waiting for it: the collective runs beside whatever the
    rank does next (the decode of its own shard does not need the other ranks' sizes).  Returns a handle for
    exchange_sizes_end."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ([int(local_size)], None, None)
    import torch
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(local_size)], dtype=torch.int64, device=device)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    work = dist.all_gather_into_tensor(sizes, mine, async_op=True)
    return (sizes, work, (dist, mine))


def exchange_sizes_end(handle) -> Tuple[List[int], int]:
    """Wait for exchange_sizes_begin's collective; returns (sizes, this rank's base offset)."""
    sizes, work, ctx = handle
    if work is None:
        return sizes, 0
    work.wait()
    lst = [int(x) for x in sizes.tolist()]
    return lst, sum(lst[:ctx[0].get_rank()])
