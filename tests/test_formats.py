"""CPU: on-disk formats of the reference (rrtqx_3d_b200/formats.py): Julia 1.0 float printing, the obstacle file
readers against the reference fixture, the dump writers against hand-built expectations."""
import io

import numpy as np
import pytest

from rrtqx_3d_b200 import formats as F
from rrtqx_3d_b200 import workloads as W


@pytest.mark.parametrize("x,s", [
    (1.0, "1.0"), (0.1, "0.1"), (-0.0, "-0.0"), (0.0, "0.0"), (1e-5, "1.0e-5"), (0.0001, "0.0001"), (0.00012345, "0.00012345"),
    (123456.789, "123456.789"), (999999.0, "999999.0"), (1e6, "1.0e6"), (1234567.0, "1.234567e6"), (1.5e10, "1.5e10"),
    (1e22, "1.0e22"), (5e-324, "5.0e-324"), (-2.5, "-2.5"), (100.0, "100.0"), (3.0e-5, "3.0e-5"), (1.0 / 3.0, "0.3333333333333333"),
    (float("inf"), "Inf"), (float("-inf"), "-Inf"), (float("nan"), "NaN"), (6.283185307179586, "6.283185307179586"),
    (1.9196069361095438, "1.9196069361095438"), (4.0e-4, "0.0004"), (-12345678.9, "-1.23456789e7")])
def test_julia_float_layout(x, s):
    # expectations follow base/grisu/grisu.jl `_show` of Julia 1.0: exponential iff pt <= -4 or pt > 6
    assert F.julia_float(x) == s


def test_julia_float_round_trips():
    rng = np.random.default_rng(3)
    for v in np.concatenate([rng.normal(size=2000) * 10.0 ** rng.integers(-12, 12, 2000), rng.random(500)]):
        assert float(F.julia_float(v)) == v


def test_str2array_and_sphere_file_round_trip(tmp_path):
    assert np.array_equal(F.str2array("1.5, -2,3e2\n"), [[1.5, -2.0, 300.0]])
    assert np.array_equal(F.str2array("7"), [[7.0]])
    centers, radii, beh = W.building2_spheres()
    p = tmp_path / "spheres.txt"
    with open(p, "w") as f:                      # the format of R/environments/building2.txt
        f.write(f"{len(radii)}\n")
        for c, r, b in zip(centers, radii, beh):
            f.write(f"{F.julia_float(c[0])}, {F.julia_float(c[1])}, {F.julia_float(c[2])}\n{F.julia_float(r)}\n{int(b)}\n")
    obs = F.read_sphere_obstacles(str(p), obs_mult=2)
    assert len(obs) == 2 * len(radii)
    assert np.array_equal(np.vstack([o["position"] for o in obs[::2]]), centers)
    assert np.array_equal([o["radius"] for o in obs[1::2]], radii)
    assert all(o["lifeSpan"] == 300.0 for o in obs)
    # building2: every obstacle is type 1 ("appearing"): senseable, unused until sensed
    assert all(o["senseableObstacle"] and o["obstacleUnused"] and not o["obstacleUnusedAfterSense"] for o in obs)
    # the same numbers through the older reader used by the workloads
    c2, r2, b2 = W.read_sphere_obstacle_file(str(p))
    assert np.array_equal(c2, centers) and np.array_equal(r2, radii) and np.array_equal(b2, beh)


def test_polygon_file_reader():
    text = "2\n4\n0, 0\n6, 0\n6, 6\n0, 6\n0\n3\n1.5, 1\n2, 4\n-1, 2\n-1\n"
    obs = F.read_polygon_obstacles(io.StringIO(text))
    assert len(obs) == 2 and obs[0]["polygon"].shape == (4, 2) and obs[1]["polygon"].shape == (3, 2)
    # Obstacle(kind, polygon): bbox midpoint and farthest-vertex radius (DRRT_data_structures.jl:229-241)
    assert np.array_equal(obs[0]["position"], [[3.0, 3.0]]) and obs[0]["radius"] == np.sqrt(18.0)
    assert obs[0]["obstacleUnused"] is False and obs[1]["obstacleUnusedAfterSense"] is True
    with pytest.raises(ValueError):
        F.read_polygon_obstacles(io.StringIO("1\n2\n0,0\n1,1\n5\n"))


def test_dump_writers():
    pos = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 3.0], [-1.5, 0.25, 1e-5], [4.0, 5.0, 6.0]])
    order = np.array([0, 2, 1, 3])                      # kd visit order
    cost = np.array([0.0, 3.5, np.inf, 7.25])
    lmc = np.array([0.0, 3.5, 2.0, 7.25])
    assert F.dump_to_string(F.save_rrt_nodes, pos, order, cost, lmc) == (
        "0.0,0.0,0.0,0.0,0.0\n-1.5,0.25,1.0e-5,Inf,2.0\n1.0,2.0,3.0,3.5,3.5\n4.0,5.0,6.0,7.25,7.25\n")
    parent = np.array([-1, 0, -1, 1])
    assert F.dump_to_string(F.save_rrt_tree, pos, order, cost, parent) == (
        "1.0,2.0,3.0,3.5\n0.0,0.0,0.0,0.0\n4.0,5.0,6.0,7.25\n1.0,2.0,3.0,3.5\n")
    row_ptr, col = np.array([0, 2, 3, 3, 4]), np.array([1, 2, 0, 1])
    assert F.dump_to_string(F.save_rrt_graph, pos, order, row_ptr, col) == (
        "0.0,0.0,0.0\n1.0,2.0,3.0\n0.0,0.0,0.0\n-1.5,0.25,1.0e-5\n1.0,2.0,3.0\n0.0,0.0,0.0\n4.0,5.0,6.0\n1.0,2.0,3.0\n")
    assert F.dump_to_string(F.save_obstacle_locations, [[1, 2, 3], [4, 5, 6]], [3.5, 1.0], [False, True]) == "1.0,2.0,3.0,3.5\n"


def test_path_kd_data_and_collision_node_writers():
    pos = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 3.0], [4.5, -1.0, 2.0], [7.0, 7.0, 7.0]])
    parent = np.array([-1, 0, 1, 2])
    out = F.dump_to_string(F.save_rrt_path_q, pos, 3, 0, parent)
    assert out == ("7.0,7.0,7.0\n4.5,-1.0,2.0\n" "4.5,-1.0,2.0\n1.0,2.0,3.0\n" "1.0,2.0,3.0\n0.0,0.0,0.0\n" "0.0,0.0,0.0\n")
    # a node whose parent edge is not in use ends the walk there (rrtParentUsed false)
    out = F.dump_to_string(F.save_rrt_path_q, pos, 3, 0, parent, parent_used=[False, True, False, True])
    assert out == "7.0,7.0,7.0\n4.5,-1.0,2.0\n4.5,-1.0,2.0\n"
    # the hop limit of the reference (i < 1000) on a parent cycle
    cyc = np.array([-1, 2, 1])
    out = F.dump_to_string(F.save_rrt_path_q, pos[:3], 1, 0, cyc)
    assert out.count("\n") == 2 * 1000 + 1
    # Dubins edges write their trajectory rows
    traj = {1: np.array([[1.0, 2.0], [0.5, 1.0], [0.0, 0.0]])}
    out = F.dump_to_string(F.save_rrt_path_q, pos, 1, 0, parent, trajectories=lambda n: traj[n])
    assert out == "1.0,2.0\n0.5,1.0\n0.0,0.0\n0.0,0.0,0.0\n"
    assert F.dump_to_string(F.save_kds_q, 0.25, 1.0e-5) == "0.25\n1.0e-5\n"
    assert F.dump_to_string(F.save_data, np.array([[1.0, 2.5], [3.0, 1e7]])) == "1.0,2.5\n3.0,1.0e7\n"
    assert F.dump_to_string(F.save_data, np.zeros((0, 3))) == ""
    order = np.array([0, 2, 1])
    out = F.dump_to_string(F.save_rrt_nodes_collision, pos[:3], order, [0.0, np.inf, 2.0], [1.0, 5.0, np.nan])
    assert out == "0.0,0.0,0.0,0.0\n4.5,-1.0,2.0,NaN\n1.0,2.0,3.0,5.0\n"


def test_julia_float_round_trips_random_bit_patterns():
    """Shortest round-trip printing: parsing the text gives back the same double, for random bit patterns (all
    exponents, subnormals included) and for values around the fixed/exponential switch points."""
    rng = np.random.default_rng(12)
    bits = rng.integers(0, 2 ** 63, 20000, dtype=np.int64).astype(np.uint64) | (rng.integers(0, 2, 20000).astype(np.uint64) << np.uint64(63))
    xs = bits.view(np.float64)
    xs = xs[np.isfinite(xs)]
    edge = np.array([1e-5, 9.999999999999999e-5, 1e-4, 1.0000000000000002e-4, 999999.9999999999, 1e6, 1000000.0000000001,
                     9999999.999999998, 1e7, 123456.7, 0.1 + 0.2, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308])
    for x in np.concatenate([xs, edge, -edge]):
        s = F.julia_float(float(x))
        assert float(s) == x, (x, s)
        assert ("e" in s) == (not (1e-4 <= abs(x) < 1e6)), (x, s)      # fixed notation exactly on [1e-4, 1e6)
        mant = s.lstrip("-").split("e")[0]
        assert "." in mant and not mant.endswith(".") and not mant.startswith("."), s
