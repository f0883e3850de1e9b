"""Host structures behave like the reference's (jlist.jl, list.jl) and the workload generator is pinned."""
import math

import numpy as np

from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.structures import (CSpace, JList, ObstacleList, RRTNode, SimpleEdge, SphereObstacle, addObsToCSpace,
                                      newEdge)


def test_jlist_is_lifo_with_keys():
    L = JList()
    assert L.pop_key() is False and L.top() is False            # jlist.jl:100-103,115-118
    for i in range(5):
        L.push(i, float(i) / 2)
    assert L.length == 5 and L.top() == 4
    assert [d for d, _ in L.items()] == [4, 3, 2, 1, 0]
    assert L.pop_key() == (4, 2.0)
    assert L.pop() == 3
    assert L.length == 3


def test_jlist_remove_middle_front_back():
    L = JList()
    nodes = [L.push(i) for i in range(5)]
    assert L.remove(nodes[2]) and [n.data for n in L] == [4, 3, 1, 0]
    assert L.remove(nodes[4]) and [n.data for n in L] == [3, 1, 0]      # front
    assert L.remove(nodes[0]) and [n.data for n in L] == [3, 1]         # back
    assert L.remove(nodes[3]) and L.remove(nodes[1]) and L.length == 0
    assert L.front is L.bound and L.back is L.bound
    assert L.remove(nodes[1]) is True                                     # empty list: true (jlist.jl:163-165)
    L.push(7)
    assert L.pop() == 7 and L.length == 0


def test_obstacle_list_pushes_at_front_and_flags():
    S = CSpace(3, -1.0, [-20] * 3, [20] * 3, [0, 0, 0], [1, 1, 1])
    a, b = SphereObstacle([0, 0, 0], 1.0), SphereObstacle([1, 1, 1], 2.0)
    addObsToCSpace(S, a)
    addObsToCSpace(S, b)
    assert list(S.obstacles) == [b, a] and S.obstacles.length == 2      # list.jl:53-59
    assert a.active()
    a.obstacleUnused = True
    assert not a.active()
    b.lifeSpan = 0.0
    assert not b.active()
    assert np.array_equal(S.width, np.full((1, 3), 40.0))


def test_node_and_edge_defaults():
    n = RRTNode([1.0, 2.0, 3.0])
    assert n.position.shape == (1, 3) and not n.kdInTree and not n.inHeap and n.rrtNeighborsOut.length == 0
    e = newEdge(n, RRTNode([0, 0, 0]))
    assert isinstance(e, SimpleEdge) and e.startNode is n


def test_rng_is_pinned():
    # splitmix64 counter-based stream: first outputs of stream 1 are fixed forever
    assert [hex(int(x)) for x in W.splitmix64(1, 0, 3)] == [hex(int(x)) for x in W.splitmix64(1, 0, 3)]
    u = W.uniform01(1, 0, 4)
    assert np.all((u >= 0) & (u < 1))
    p = W.uniform_points(1, 2, [-20] * 3, [20] * 3)
    assert p.tobytes().hex()[:16] == W.uniform_points(1, 1, [-20] * 3, [20] * 3).tobytes().hex()[:16]
    # lo + u .* width, in that order
    assert p[0, 0] == -20.0 + u[0] * 40.0
    # independent of how the stream is chunked
    assert np.array_equal(W.uniform01(5, 10, 5), W.uniform01(5, 0, 15)[10:])


def test_sphere_file_parser(tmp_path):
    f = tmp_path / "obs.txt"
    f.write_text("2\n-14.0, -14.0, -18.0\n3.5\n1\n1.5, 2.5, 3.5\n0.25\n0\n")
    c, r, b = W.read_sphere_obstacle_file(str(f))
    assert c.tolist() == [[-14.0, -14.0, -18.0], [1.5, 2.5, 3.5]] and r.tolist() == [3.5, 0.25] and b.tolist() == [1, 0]
