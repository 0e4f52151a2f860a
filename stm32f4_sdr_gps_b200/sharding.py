"""Sharding of the hot path over the GPUs of one box.

Units are independent - satellites for tracking, (satellite, Doppler bin) cells for acquisition - so
there is no data-path collective: every rank works on its own slice of the satellite list against its
own copy of the signal (2 MB per second of signal; replicating it is cheaper than any broadcast).  The
one exchange is the gather of the sweep triples (8 bytes per cell) so that every rank can run the host
votes on the complete Doppler x satellite grid.  `torch.distributed` is plumbing here: NCCL over
NVLink for device tensors, gloo for the CPU tests.
"""
from __future__ import annotations

import numpy as np

SEARCH_RES_WORDS = 4  # max, phase, avg, reserved (uint16 each), include/gpsb.h gpsb_search_res


def shard_satellites(n_sv: int, rank: int, world: int) -> np.ndarray:
    """Indices (into the caller's satellite list) owned by `rank`: round-robin, so that a sorted PRN list
    with a few strong low PRNs does not pile up on rank 0."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return np.arange(rank, n_sv, world, dtype=np.int64)


def owner_of(sv_index: int, world: int) -> int:
    return sv_index % world


def padded_count(n_sv: int, world: int) -> int:
    """Every rank contributes the same number of rows to the gather (all_gather needs equal shapes)."""
    return (n_sv + world - 1) // world


def pack_local(res: np.ndarray, n_sv: int, rank: int, world: int) -> np.ndarray:
    """res: this rank's sweep result, shape (n_local, n_bins, n_ms) of gpsb_search_res (or (.., 4) uint16).
    Returns a (padded_count, n_bins, n_ms, 4) int16 block; rows beyond n_local are zero."""
    mine = shard_satellites(n_sv, rank, world)
    raw = np.ascontiguousarray(res).view(np.uint16).reshape(len(mine), *res.shape[1:3], SEARCH_RES_WORDS)
    out = np.zeros((padded_count(n_sv, world),) + raw.shape[1:], np.uint16)
    out[:len(mine)] = raw
    return out.view(np.int16)


def unpack_gathered(blocks, n_sv: int, world: int) -> np.ndarray:
    """blocks[r] = rank r's pack_local() block.  Returns the full (n_sv, n_bins, n_ms, 4) uint16 grid in the
    caller's satellite order."""
    first = np.asarray(blocks[0])
    full = np.zeros((n_sv,) + first.shape[1:], np.uint16)
    for r in range(world):
        mine = shard_satellites(n_sv, r, world)
        full[mine] = np.asarray(blocks[r]).view(np.uint16)[:len(mine)]
    return full


def gather_sweep(local_block, n_sv: int, world: int, device=None) -> np.ndarray:
    """All-gather the per-rank blocks (torch.distributed must be initialised when world > 1).
    `local_block` is pack_local()'s array; with `device` set the exchange runs on device tensors (NCCL)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return unpack_gathered([local_block], n_sv, 1)
    # exchanged as int32 pairs: gloo has no 16-bit integer type, and 8-byte triples are 2 words anyway
    words = np.ascontiguousarray(local_block).view(np.int32)
    t = torch.from_numpy(words)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return unpack_gathered([p.cpu().numpy().view(np.int16) for p in parts], n_sv, world)


def reduce_max_time(seconds: float, world: int, device=None) -> float:
    """Job time = slowest rank (device-timed per rank, max over ranks)."""
    if world == 1:
        return float(seconds)
    import torch
    import torch.distributed as dist

    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- acquisition, round 2: the sweep is sharded by (bin, ms) CELL GROUP, not by satellite (gpsb_sweep_gather) ----------
# A cell group is one mixed millisecond serving every satellite searched on it, so a rank that owns whole groups keeps
# whole 8-satellite tiles of the dp4a search whatever the number of ranks is.  These are the index rules of the C ABI
# (csrc/gpsb_cuda.cu: gpsb_sweep_gather_dev, k_unshard), restated for the CPU tests of the N > 1 path.
def group_owner(group: int, world: int) -> int:
    """Rank that computes cell group `group` = bin * n_ms + ms: round-robin."""
    return group % world


def groups_of(rank: int, n_groups: int, world: int) -> np.ndarray:
    return np.arange(rank, n_groups, world, dtype=np.int64)


def group_block(full: np.ndarray, rank: int, world: int) -> np.ndarray:
    """This rank's dense block of a sweep `full` (n_sv, n_bins, n_ms, 4): rows = its k-th group, padded to the common
    block size ceil(groups / world); block[k, v] = the triple of satellite v on group rank + k * world."""
    n_sv, n_bins, n_ms = full.shape[:3]
    n_groups = n_bins * n_ms
    n_local = (n_groups + world - 1) // world
    flat = full.reshape(n_sv, n_groups, -1)
    block = np.zeros((n_local, n_sv, flat.shape[-1]), full.dtype)
    mine = groups_of(rank, n_groups, world)
    block[:len(mine)] = flat[:, mine].transpose(1, 0, 2)
    return block


def unshard_groups(blocks, n_sv: int, n_bins: int, n_ms: int) -> np.ndarray:
    """The all-gathered blocks (blocks[r] = rank r's group_block) -> the (n_sv, n_bins, n_ms, 4) grid: k_unshard."""
    world = len(blocks)
    first = np.asarray(blocks[0])
    grid = np.zeros((n_sv, n_bins * n_ms, first.shape[-1]), first.dtype)
    for g in range(n_bins * n_ms):
        grid[:, g] = np.asarray(blocks[g % world])[g // world]
    return grid.reshape(n_sv, n_bins, n_ms, -1)


def gather_groups(local_block: np.ndarray, world: int):
    """All-gather of the per-rank group blocks over torch.distributed (gloo in the CPU tests; the product exchanges them
    with ncclAllGather inside gpsb_sweep_gather).  Returns the list of blocks, rank order."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local_block]
    t = torch.from_numpy(np.ascontiguousarray(local_block).view(np.int16).astype(np.int32))   # gloo has no 16-bit integers
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return [p.numpy().astype(np.int16).view(np.uint16).reshape(local_block.shape) for p in parts]
