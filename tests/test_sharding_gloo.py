"""N > 1 path on CPU: two gloo ranks shard bunches, take disjoint RNG slices, and the merge by
identifier reproduces the single-rank bookkeeping.  The propagation itself is stood in for by the
oracle on a tiny scene (there is no GPU here); what is tested is the host-side sharding logic
bench.py and a multi-converter caller use."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clsim_b200 import capi, steps
from clsim_b200.sharding import RNG_ROWS_PER_DEVICE, merge_results, rng_row_offset, shard_bunches, split_steps


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle
    from tests.scenes import make_scene
    sc = make_scene("homogeneous", geo_kind="ring")
    osc = pyoracle.Scene(sc.medium, sc.geo, sc.generators, sc.bias, sc.options())
    num_bunches = 6
    mine = shard_bunches(num_bunches, rank, world)
    a = capi.safeprime_multipliers(rng_row_offset(rank), 64)
    x = pyoracle.seed_states(1000 + rank, a)
    photons_total, hits_total = 0, 0
    results = []
    for b in mine:
        src = (sc.geo.posX[b] + 4.0, sc.geo.posY[b], sc.geo.posZ[b])
        bunch = steps.point_source_steps(64, 40, pos=src, seed=b)
        bunch["identifier"] = b
        ph, cnt, st, x, _ = osc.propagate(bunch, x, a)
        results.append((b, ph))
        photons_total += st["photons"]
        hits_total += cnt
    t = torch.tensor([photons_total, hits_total, float(a[0]), float(a[-1])], dtype=torch.float64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    tmax = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)   # the max-over-ranks timing reduction bench.py does
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([g.numpy() for g in gathered]))
    np.save(os.path.join(out_dir, "ids%d.npy" % rank), np.array([i for i, _ in results]))
    np.save(os.path.join(out_dir, "tmax%d.npy" % rank), tmax.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_merge(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0 = np.load(tmp_path / "rank0.npy")
    g1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(g0, g1)                       # both ranks see the same gathered totals
    assert g0[:, 0].sum() == 6 * 64 * 40                # every photon of every bunch exactly once
    ids = sorted(list(np.load(tmp_path / "ids0.npy")) + list(np.load(tmp_path / "ids1.npy")))
    assert ids == list(range(6))
    assert np.load(tmp_path / "tmax0.npy")[0] == pytest.approx(0.2) == np.load(tmp_path / "tmax1.npy")[0]
    # RNG slices of the two ranks are disjoint rows of the descending multiplier sequence
    assert g0[0, 3] > g0[1, 2]


def test_sharding_helpers():
    assert rng_row_offset(0) == 0 and rng_row_offset(3) == 3 * RNG_ROWS_PER_DEVICE
    assert RNG_ROWS_PER_DEVICE >= 2 * 148 * 1024
    assert shard_bunches(7, 1, 3) == [1, 4]
    s = steps.muon_track_steps(1000, seed=1)
    parts = split_steps(s, 4, granularity=64)
    assert sum(len(p) for p in parts) == 1000
    assert all(len(p) % 64 == 0 for p in parts[:-1])
    merged = merge_results([(1, s[:10]), (2, s[10:20]), (1, s[20:25])])
    assert len(merged[1]) == 15 and len(merged[2]) == 10
    a0 = capi.safeprime_multipliers(rng_row_offset(0), 16)
    a1 = capi.safeprime_multipliers(rng_row_offset(1), 16)
    assert a1.max() < a0.min()
