"""Multi-GPU entry points of the C ABI (rrtqx_comm_*, rrtqx_edge_check_batch_sharded): the one-rank communicator on
any box, the ONE-process / N-context form and the torchrun form on boxes with at least two GPUs (they skip with
fewer).  Gathered flags and counts are compared with the oracle on every rank."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Comm, Context, DeviceTree, SphereSet, edge_check_batch, unpack_flags

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _workload(n_nodes=20000, n_edges=300001):
    pts, qs, r = W.c2_workload(n_nodes, 4096)
    u = W.splitmix64(9, 0, 2 * n_edges)
    src = (u[:n_edges] % np.uint64(n_nodes)).astype(np.int32)
    dst = ((src.astype(np.int64) + 1 + (u[n_edges:] % np.uint64(50)).astype(np.int64)) % n_nodes).astype(np.int32)
    dst[::101] = src[::101]                       # zero-length edges
    c, rad = W.c3_obstacles(64)
    return pts, qs, r, src, dst, c, rad


def _oracle_flags(pts, src, dst, c, rad):
    sph, ns = oracle.make_spheres(c, rad)
    out = np.zeros(len(src), dtype=np.uint8)
    pts = np.ascontiguousarray(pts)
    oracle.lib().orc_edge_check_batch(sph, ns, oracle._p(pts, oracle.c_f64p), 3, oracle._p(src, oracle.c_i32p),
                                      oracle._p(dst, oracle.c_i32p), 0, len(src), W.ROBOT_RADIUS, 0,
                                      oracle._p(out, oracle.c_u8p), 4)
    return out


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_one_rank_communicator_packs_the_same_flags(ctx):
    import torch
    pts, qs, r, src, dst, c, rad = _workload()
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    S = SphereSet(ctx, c, rad)
    comm = Comm.local([ctx])
    info = comm.info()
    assert info["n_ranks"] == 1 and info["n_local"] == 1 and not info["peer_stores"]
    total, per = comm.packed_words(len(src))
    assert total == per == (len(src) + 31) // 32
    ds, dd = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
    out = torch.zeros(total, dtype=torch.int32, device="cuda")
    comm.edge_check_sharded([t], [S], [ds], [dd], len(src), W.ROBOT_RADIUS, [out])
    got = unpack_flags(out.cpu().numpy().view(np.uint32), len(src))
    want = _oracle_flags(pts, src, dst, c, rad)
    assert np.array_equal(got, want) and np.array_equal(got, edge_check_batch(t, S, src, dst, W.ROBOT_RADIUS))
    # all-gather of one rank = a copy
    a = torch.arange(1000, dtype=torch.int32, device="cuda")
    b = torch.zeros_like(a)
    torch.cuda.synchronize()
    comm.allgather([a], [b], a.numel() * 4)
    ctx.sync()
    assert torch.equal(a, b)
    comm.close()


@pytest.mark.skipif("_n_gpus() < 2")
def test_one_process_two_contexts_sharded_check_and_range_queries():
    """The form a Julia host uses: one process, one context per device.  Also the regression test of the per-device
    shared-memory opt-in of the range kernel (a second context on another device used to launch without it)."""
    import torch
    n = min(_n_gpus(), 4)
    pts, qs, r, src, dst, c, rad = _workload()
    want = _oracle_flags(pts, src, dst, c, rad)
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    oc, _, _, _ = orc.range_batch(r, qs, want_lists=False, nthreads=8)
    ctxs = [Context(g) for g in range(n)]
    trees, sph, dsrc, ddst, outs = [], [], [], [], []
    comm = Comm.local(ctxs)
    info = comm.info()
    assert info["n_ranks"] == n and info["n_local"] == n
    total, per = comm.packed_words(len(src))
    for g, cx in enumerate(ctxs):
        t = DeviceTree(cx, 3)
        t.insert_batch(pts)
        trees.append(t)
        sph.append(SphereSet(cx, c, rad))
        dev = torch.device("cuda", g)
        dsrc.append(torch.from_numpy(src).to(dev))
        ddst.append(torch.from_numpy(dst).to(dev))
        outs.append(torch.zeros(total, dtype=torch.int32, device=dev))
    for g in range(n):
        torch.cuda.synchronize(g)
    for rep in range(3):
        comm.edge_check_sharded(trees, sph, dsrc, ddst, len(src), W.ROBOT_RADIUS, outs)
        for g in range(n):
            got = unpack_flags(outs[g].cpu().numpy().view(np.uint32), len(src))
            assert np.array_equal(got, want), (rep, g, info)
            outs[g].zero_()
        for g in range(n):
            torch.cuda.synchronize(g)
    # sharded range queries: rank g owns a slice of the batch, the per-query counts are all-gathered
    lo = [len(qs) * g // n for g in range(n + 1)]
    per_q = max(lo[g + 1] - lo[g] for g in range(n))
    send, recv, results = [], [], []
    for g, t in enumerate(trees):
        res, _ = t.range_query(qs[lo[g]:lo[g + 1]], r)
        results.append(res)
        dev = torch.device("cuda", g)
        buf = torch.zeros(per_q, dtype=torch.int32, device=dev)
        cnt, _ = res.layout()
        buf[:len(cnt)] = torch.from_numpy(cnt).to(dev)
        send.append(buf)
        recv.append(torch.zeros(per_q * n, dtype=torch.int32, device=dev))
    for g in range(n):
        torch.cuda.synchronize(g)
    comm.allgather(send, recv, per_q * 4, side_stream=True)
    comm.join()
    for cx in ctxs:
        cx.sync()
    for g in range(n):
        full = recv[g].cpu().numpy().reshape(n, per_q)
        got = np.concatenate([full[k, :lo[k + 1] - lo[k]] for k in range(n)])
        assert np.array_equal(got, oc), g
    comm.close()


@pytest.mark.skipif("_n_gpus() < 2")
def test_torchrun_two_ranks_sharded_check_matches_oracle(tmp_path):
    """One process per GPU: unique id from rank 0 moved with torch.distributed, NCCL gather inside the library."""
    out = tmp_path / "mp.log"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tests", "mp_sharded_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out.write_text(p.stdout + p.stderr)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert p.stdout.count("rank ok") == 2, p.stdout[-2000:]
