"""Multi-rank path on CPU: world_size-2 gloo processes shard a seeded episode batch, solve their slices independently
(through the host build of the kernel's solver text) and reduce run statistics.  The union of the shards must equal the
unsharded run bit for bit, and the reductions must match."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from tests.helpers import PKG, ROOT
from _bmpc import scenarios, shard

B = 24


def _worker(rank, world_size, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world_size), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from tests.hostsim.driver import HostSim
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=77)
    lo, hi = shard.shard_bounds(B, world_size, rank)
    hs = HostSim(scenarios.highway_config(), hi - lo)
    r = hs.solve(x0[lo:hi], z0[lo:hi], xref[lo:hi], pp[lo:hi])
    stats = shard.reduce_stats({"solves": hi - lo, "sum_iters": int(r["iters"].sum()), "max_iters": int(r["iters"].max()),
                                "polished": int((r["status"] == 0).sum())})
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), u0=r["u0"], objective=r["objective"], lo=lo, hi=hi,
             **{"stat_" + k: v for k, v in stats.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from tests.hostsim.driver import HostSim
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=77)
    ref = HostSim(scenarios.highway_config(), B).solve(x0, z0, xref, pp)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(2)]
    assert [int(p["lo"]) for p in parts] == [0, 12] and [int(p["hi"]) for p in parts] == [12, 24]
    assert np.array_equal(np.vstack([p["u0"] for p in parts]), ref["u0"])
    assert np.array_equal(np.concatenate([p["objective"] for p in parts]), ref["objective"])
    for p in parts:     # every rank holds the same reduced statistics
        assert float(p["stat_solves"]) == B
        assert float(p["stat_sum_iters"]) == float(ref["iters"].sum())
        assert float(p["stat_max_iters"]) == float(ref["iters"].max())
        assert float(p["stat_polished"]) == float((ref["status"] == 0).sum())


def test_shard_bounds_cover_every_episode_once():
    for total in (0, 1, 7, 16384, 65536):
        for ws in (1, 2, 3, 8):
            cuts = [shard.shard_bounds(total, ws, r) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
