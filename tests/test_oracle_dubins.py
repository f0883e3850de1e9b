"""CPU: pins the oracle's Dubins restatement (DRRT_DubinsEdge_functions.jl:329-709, :70-95) with hand-evaluated
cases, the textbook CSC closed forms (an independent derivation) and geometric invariants of the trajectory."""
import math

import numpy as np

import oracle

TWO_PI = 2.0 * math.pi


def _mod2pi(a):
    return a - TWO_PI * math.floor(a / TWO_PI)


def _csc_textbook(s, g, r):
    """Shkel & Lumelsky closed forms of the four CSC words (lengths); None where the word does not exist."""
    dx, dy = g[0] - s[0], g[1] - s[1]
    d = math.hypot(dx, dy) / r
    th = math.atan2(dy, dx)
    a, b = _mod2pi(s[3] - th), _mod2pi(g[3] - th)
    sa, sb, ca, cb, cab = math.sin(a), math.sin(b), math.cos(a), math.cos(b), math.cos(a - b)
    out = {}
    p2 = 2 + d * d - 2 * cab + 2 * d * (sa - sb)
    if p2 >= 0:
        t = math.atan2(cb - ca, d + sa - sb)
        out["lsl"] = (_mod2pi(-a + t) + math.sqrt(p2) + _mod2pi(b - t)) * r
    p2 = 2 + d * d - 2 * cab + 2 * d * (sb - sa)
    if p2 >= 0:
        t = math.atan2(ca - cb, d - sa + sb)
        out["rsr"] = (_mod2pi(a - t) + math.sqrt(p2) + _mod2pi(-b + t)) * r
    p2 = -2 + d * d + 2 * cab + 2 * d * (sa + sb)
    if p2 >= 0:
        p = math.sqrt(p2)
        t = math.atan2(-ca - cb, d + sa + sb) - math.atan2(-2.0, p)
        out["lsr"] = (_mod2pi(-a + t) + p + _mod2pi(-_mod2pi(b) + t)) * r
    p2 = d * d - 2 + 2 * cab - 2 * d * (sa + sb)
    if p2 >= 0:
        p = math.sqrt(p2)
        t = math.atan2(ca + cb, d - sa - sb) - math.atan2(2.0, p)
        out["rsl"] = (_mod2pi(a - t) + p + _mod2pi(b - t)) * r
    return out


def test_straight_ahead_is_the_degenerate_rsl_with_four_rows():
    # both headings along +x, r = 1: irc = (0,-1), glc = (10,1), v = (10,2)/sqrt(104), R = -2/sqrt(104),
    # sqrt(1-R^2) = 10/sqrt(104)  =>  a = R v1 + v2 sq = 0, b = R v2 - v1 sq = -1: the inner tangent points are
    # the start and the goal themselves, rsl = 0 + 10 + 0.  rsl is tried first and the later ties (rsr, lsl = 10)
    # lose to the strict `bestDist > length` (DRRT_DubinsEdge_functions.jl:384-387, 404-407).
    dist, typ, traj = oracle.dubins_trajectory([0, 0, 0, 0], [10, 0, 0, 0], 1.0)
    assert typ == 0 and abs(dist - 10.0) < 1e-12
    assert traj.shape == (4, 2)
    assert np.allclose(traj, [[0, 0], [0, 0], [10, 0], [10, 0]], atol=1e-12)


def test_quarter_left_turns_known_length():
    # start heading +x at the origin, goal heading +y at (r + 5, r + 5) with r = 1: one quarter left turn cannot
    # reach it, lsl = quarter turn + straight 5*sqrt(2)... evaluated against the textbook closed form instead
    s, g, r = [0.0, 0.0, 0.0, 0.0], [6.0, 6.0, 0.0, math.pi / 2], 1.0
    dist, typ, traj = oracle.dubins_trajectory(s, g, r)
    ref = _csc_textbook(s, g, r)
    assert abs(dist - min(ref.values())) < 1e-9
    assert oracle.lib() is not None and typ in (0, 1, 3, 4)


def test_csc_lengths_agree_with_textbook_closed_forms():
    rng = np.random.default_rng(5)
    names = ["rsl", "rsr", "rlr", "lsr", "lsl", "lrl"]
    n_checked = 0
    for _ in range(4000):
        r = rng.uniform(0.5, 2.0)
        s = [rng.uniform(-50, 50), rng.uniform(-50, 50), 0.0, rng.uniform(0, TWO_PI)]
        ang, rad = rng.uniform(0, TWO_PI), rng.uniform(4.5 * r, 30.0)   # far apart: a CSC word is optimal
        g = [s[0] + rad * math.cos(ang), s[1] + rad * math.sin(ang), 0.0, rng.uniform(0, TWO_PI)]
        dist, typ, traj = oracle.dubins_trajectory(s, g, r)
        ref = _csc_textbook(s, g, r)
        best = min(ref, key=ref.get)
        assert abs(dist - ref[best]) <= 1e-9 * max(1.0, dist), (s, g, r, dist, ref)
        if abs(sorted(ref.values())[1] - ref[best]) > 1e-6:   # unambiguous winner: same word
            assert names[typ] == best
        n_checked += 1
    assert n_checked == 4000


def test_trajectory_invariants():
    rng = np.random.default_rng(6)
    for _ in range(3000):
        r = 1.0
        s = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), 0.0, rng.uniform(0, TWO_PI)])
        ang, rad = rng.uniform(0, TWO_PI), rng.uniform(0.05, 8.0)   # includes the D < 4r regime (CCC words)
        g = np.array([s[0] + rad * math.cos(ang), s[1] + rad * math.sin(ang), 0.0, rng.uniform(0, TWO_PI)])
        dist, typ, traj = oracle.dubins_trajectory(s, g, r)
        assert 0 <= typ <= 5 and np.isfinite(dist)
        assert dist >= math.hypot(*(g[:2] - s[:2])) - 1e-9
        # starts on the start location; ends within one arc step (0.1 rad * r) of the goal
        assert np.allclose(traj[0], s[:2], atol=1e-9)
        assert math.hypot(*(traj[-1] - g[:2])) <= 0.1 * r + 1e-9
        # every arc sample lies on a circle of radius r: consecutive samples are at most one chord apart,
        # except the straight part; polyline length <= word length and within the chord error of it
        seg = np.hypot(*np.diff(traj, axis=0).T)
        poly = seg.sum() + math.hypot(*(traj[-1] - g[:2]))
        assert poly <= dist + 1e-9
        assert poly >= dist * (1.0 - 0.1 ** 2 / 24.0) - 0.3 * r   # chord/arc ratio at 0.1 rad, three partial steps
        assert len(traj) <= 3 * 64 + 2


def test_mirror_symmetry_swaps_left_and_right():
    rng = np.random.default_rng(8)
    swap = {0: 3, 3: 0, 1: 4, 4: 1, 2: 5, 5: 2}
    for _ in range(500):
        s = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), 0.0, rng.uniform(0.1, TWO_PI - 0.1)])
        g = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), 0.0, rng.uniform(0.1, TWO_PI - 0.1)])
        d1, t1, _ = oracle.dubins_trajectory(s, g, 1.0)
        sm, gm = s * [1, -1, 1, 1], g * [1, -1, 1, 1]
        sm[3], gm[3] = TWO_PI - s[3], TWO_PI - g[3]
        d2, t2, _ = oracle.dubins_trajectory(sm, gm, 1.0)
        assert abs(d1 - d2) < 1e-9
        if t2 != swap[t1]:   # only under a length tie between words
            assert abs(d1 - d2) < 1e-9


def test_saturate_dubins_matches_hand_evaluation():
    # within delta: untouched
    p = oracle.saturate_dubins([1.0, 0.0, 0.0, 0.5], [0.0, 0.0, 0.0, 0.4], 5.0)
    assert np.array_equal(p, [1.0, 0.0, 0.0, 0.5])
    # beyond delta, small heading difference: all four coordinates scaled by delta / R3SDist
    new, c, delta = np.array([6.0, 8.0, 0.0, 1.0]), np.array([0.0, 0.0, 0.0, 0.5]), 5.0
    d = math.sqrt(6.0 ** 2 + 8.0 ** 2 + 0.5 ** 2)
    p = oracle.saturate_dubins(new, c, delta)
    assert np.allclose(p, c + (new - c) * delta / d, rtol=1e-15)
    # heading difference >= pi: the new heading is unwrapped first, then clamped into [0, 2 pi]
    new, c = np.array([3.0, 4.0, 0.0, 6.0]), np.array([0.0, 0.0, 0.0, 0.2])
    wrap = min(abs(6.0 - 0.2), 0.2 + TWO_PI - 6.0)
    d = math.sqrt(25.0 + wrap ** 2)
    p = oracle.saturate_dubins(new, c, 2.0)
    th = 0.2 + ((6.0 - TWO_PI) - 0.2) * 2.0 / d
    assert np.allclose(p[:3], (new[:3] - c[:3]) * 2.0 / d + c[:3], rtol=1e-15)
    assert abs(p[3] - max(min(th, TWO_PI), 0.0)) < 1e-15
