"""Multi-process (gloo, world_size 2) test of the batch sharding + fixed-size result gather that
bench.py and the multi-GPU path use; the per-shard work is done by the oracle here (CPU)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from rrtqx_3d_b200.sharding import shard_bounds, shard_sizes


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 1000, 10 ** 7 + 3):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, g, w) for g in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert sum(shard_sizes(n, w)) == n and max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import oracle
    from rrtqx_3d_b200 import workloads as W
    from rrtqx_3d_b200.sharding import gather_fixed, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts, qs, r = W.c2_workload(5000, 301)           # replicated tree, odd batch size
    t = oracle.KDTree(3)
    t.insert_batch(pts)
    lo, hi = shard_bounds(len(qs), rank, world)
    counts, _, _, _ = t.range_batch(r, qs[lo:hi], want_lists=False)
    idx, dd = t.nearest_batch(qs[lo:hi])
    full_counts = gather_fixed(counts, len(qs), dist)
    full_nn = gather_fixed(idx, len(qs), dist)
    full_dd = gather_fixed(dd, len(qs), dist)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), counts=full_counts, nn=full_nn, dd=full_dd)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_equals_unsharded(tmp_path):
    import oracle
    from rrtqx_3d_b200 import workloads as W
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    pts, qs, r = W.c2_workload(5000, 301)
    t = oracle.KDTree(3)
    t.insert_batch(pts)
    counts, _, _, _ = t.range_batch(r, qs, want_lists=False)
    idx, dd = t.nearest_batch(qs)
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{rank}.npz"))
        assert np.array_equal(z["counts"], counts)
        assert np.array_equal(z["nn"], idx)
        assert np.array_equal(z["dd"].view(np.uint64), dd.view(np.uint64))
