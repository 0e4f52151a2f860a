"""Multi-GPU host logic on CPU: 2-process gloo world.  Each rank owns a round-robin slice of the
satellite list, produces its sweep block and all-gathers; every rank must end up with the same complete
grid and the same Doppler votes as a single-process run."""
import os
import socket

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_cells(n_sv, n_bins, n_ms):
    """Deterministic stand-in for sweep triples (the parity of real cells is tested on the GPU)."""
    sv, b, m = np.meshgrid(np.arange(n_sv), np.arange(n_bins), np.arange(n_ms), indexing="ij")
    out = np.zeros((n_sv, n_bins, n_ms, 4), np.uint16)
    out[..., 0] = 200 + (sv * 7 + b * 3 + m) % 400
    out[..., 1] = (sv * 131 + b * 17 + m * 5) % 2046
    out[..., 2] = 40 + (sv + b) % 5
    return out


def _worker(rank, world, port, n_sv, n_bins, n_ms, q):
    try:
        _worker_body(rank, world, port, n_sv, n_bins, n_ms, q)
    except Exception as e:      # surface the failure instead of letting the parent time out
        q.put((rank, False, repr(e), []))


def _worker_body(rank, world, port, n_sv, n_bins, n_ms, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _fake_cells(n_sv, n_bins, n_ms)
    mine = sharding.shard_satellites(n_sv, rank, world)
    block = sharding.pack_local(full[mine], n_sv, rank, world)
    got = sharding.gather_sweep(block, n_sv, world)
    t = sharding.reduce_max_time(0.5 + rank, world)
    q.put((rank, bool(np.array_equal(got, full)), t, mine.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_sv", [32, 5, 1])
def test_two_rank_gloo_gather(n_sv):
    import torch.multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_sv, 21, 10, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _, _ in res), res
    assert all(abs(t - 1.5) < 1e-12 for _, _, t, _ in res)              # max over ranks
    owned = sorted(res[0][3] + res[1][3])
    assert owned == list(range(n_sv))                                   # a partition: nothing lost, nothing twice


def test_partition_properties():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 4, 31, 32, 33):
            parts = [sharding.shard_satellites(n, r, world) for r in range(world)]
            allv = np.concatenate(parts) if parts else np.zeros(0)
            assert sorted(allv.tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert all(sharding.owner_of(int(i), world) == r for r, p in enumerate(parts) for i in p)
            assert sharding.padded_count(n, world) == max([len(p) for p in parts] + [0])
    with pytest.raises(ValueError):
        sharding.shard_satellites(4, 2, 2)


def test_single_rank_roundtrip():
    full = _fake_cells(7, 3, 2)
    block = sharding.pack_local(full, 7, 0, 1)
    assert np.array_equal(sharding.gather_sweep(block, 7, 1), full)


# ---- round 2: the sweep sharded by (bin, ms) cell group (gpsb_sweep_gather), index rules on the CPU ---------------------
def _group_worker(rank, world, port, n_sv, n_bins, n_ms, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        full = _fake_cells(n_sv, n_bins, n_ms)
        block = sharding.group_block(full, rank, world)           # what this rank's sharded sweep writes (dense block)
        got = sharding.unshard_groups(sharding.gather_groups(block, world), n_sv, n_bins, n_ms)
        q.put((rank, bool(np.array_equal(got, full)), int(len(sharding.groups_of(rank, n_bins * n_ms, world)))))
        dist.destroy_process_group()
    except Exception as e:
        q.put((rank, False, repr(e)))


@pytest.mark.parametrize("shape", [(32, 21, 10), (5, 29, 10), (3, 1, 1)])
def test_two_rank_gloo_group_sharded_sweep(shape):
    """Two ranks each compute their share of the (bin, ms) cell groups, all-gather the dense blocks and permute them
    into the (sv, bin, ms) grid: every rank ends up with the whole sweep.  (21 x 10 = 210 groups -> 105 each; 1 group
    -> one rank idle with a padded block.)"""
    import torch.multiprocessing as mp
    n_sv, n_bins, n_ms = shape
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_group_worker, args=(r, 2, port, n_sv, n_bins, n_ms, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] is True for r in res), res
    assert res[0][2] + res[1][2] == n_bins * n_ms and abs(res[0][2] - res[1][2]) <= 1


def test_group_sharding_rules():
    for world in (1, 2, 3, 4, 8):
        full = _fake_cells(9, 7, 5)
        blocks = [sharding.group_block(full, r, world) for r in range(world)]
        assert len({b.shape for b in blocks}) == 1                                   # equal blocks: all_gather needs that
        assert np.array_equal(sharding.unshard_groups(blocks, 9, 7, 5), full)
        owned = np.concatenate([sharding.groups_of(r, 35, world) for r in range(world)])
        assert sorted(owned.tolist()) == list(range(35))
        assert all(sharding.group_owner(int(g), world) == r for r in range(world) for g in sharding.groups_of(r, 35, world))
