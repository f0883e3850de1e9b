"""torchrun worker of tests/test_gpu_multi.py: every rank holds the replicated tree / obstacles / edge list, the
library shards the check and gathers the packed flags over NCCL; every rank compares ALL flags with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from rrtqx_3d_b200 import workloads as W  # noqa: E402
from rrtqx_3d_b200.device import Comm, Context, DeviceTree, SphereSet, unpack_flags  # noqa: E402
from test_gpu_multi import _oracle_flags, _workload  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local, torch.cuda.current_stream().cuda_stream)
    uid = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = Comm.rank(ctx, uid[0], rank, world)
    info = comm.info()
    assert info["n_ranks"] == world and info["n_local"] == 1 and info["first_rank"] == rank
    pts, qs, r, src, dst, c, rad = _workload()
    want = _oracle_flags(pts, src, dst, c, rad)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    S = SphereSet(ctx, c, rad)
    total, per = comm.packed_words(len(src))
    ds, dd = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    out = torch.zeros(total, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for rep in range(3):
        comm.edge_check_sharded([t], [S], [ds], [dd], len(src), W.ROBOT_RADIUS, [out])
        got = unpack_flags(out.cpu().numpy().view(np.uint32), len(src))
        assert np.array_equal(got, want), (rank, rep)
        out.zero_()
        torch.cuda.synchronize()
        dist.barrier()
    # per-query counts of sharded range queries, gathered by the library on its side stream
    lo = [len(qs) * g // world for g in range(world + 1)]
    per_q = max(lo[g + 1] - lo[g] for g in range(world))
    res, _ = t.range_query(qs[lo[rank]:lo[rank + 1]], r)
    cnt, _ = res.layout()
    send = torch.zeros(per_q, dtype=torch.int32, device="cuda")
    send[:len(cnt)] = torch.from_numpy(cnt).cuda()
    recv = torch.zeros(per_q * world, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    comm.allgather([send], [recv], per_q * 4, side_stream=True)
    comm.join()
    ctx.sync()
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    oc, _, _, _ = orc.range_batch(r, qs, want_lists=False, nthreads=4)
    full = recv.cpu().numpy().reshape(world, per_q)
    got = np.concatenate([full[k, :lo[k + 1] - lo[k]] for k in range(world)])
    assert np.array_equal(got, oc)
    comm.close()
    dist.barrier()
    print(f"rank ok {rank} nccl {info['nccl_version']}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
