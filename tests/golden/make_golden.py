"""Regenerates the committed fixtures under tests/golden/ (run in the authoring
container, where /root/reference exists; the GPU box only reads the outputs).

  building2_spheres.csv  <- reference fixture environments/building2.txt
                            (31 spheres; x,y,z,radius,behaviour per line), parsed
                            with the format rules of readDiscoverable3DObstaclesFromfile
                            (DRRT_Q.jl:901-947).
  oracle_vectors.json    <- seeded inputs and the oracle's outputs for every hot-path
                            primitive; lets the GPU box detect an oracle that was
                            built differently (compiler / flags) from the one that
                            was validated here.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from rrtqx_3d_b200 import workloads as W  # noqa: E402
import oracle  # noqa: E402

REF = "/root/reference/code_RRTQx_3D"


def main():
    c, rad, b = W.read_sphere_obstacle_file(os.path.join(REF, "environments", "building2.txt"))
    with open(os.path.join(HERE, "building2_spheres.csv"), "w") as f:
        f.write("# x,y,z,radius,behaviour  (from reference environments/building2.txt)\n")
        for i in range(len(rad)):
            f.write(f"{float(c[i,0])!r},{float(c[i,1])!r},{float(c[i,2])!r},{float(rad[i])!r},{int(b[i])}\n")

    L = oracle.lib()
    vec = {}
    # metrics
    pts = W.uniform_points(11, 64, [-20] * 4, [20, 20, 20, 6.283185307179586])
    vec["euclid3"] = [float(L.orc_euclid(oracle._p(pts[i], oracle.c_f64p), oracle._p(pts[i + 1], oracle.c_f64p), 3)).hex()
                      for i in range(0, 32)]
    vec["r3sdist"] = [float(L.orc_r3sdist(oracle._p(pts[i], oracle.c_f64p), oracle._p(pts[i + 1], oracle.c_f64p))).hex()
                      for i in range(0, 32)]
    # kd tree: topology + range + nearest on a 2000-node tree
    nodes, qs, _ = W.c2_workload(2000, 40)
    t = oracle.KDTree(3)
    t.insert_batch(nodes)
    par, cl, cr, sp = t.fields()
    vec["kd_parent_crc"] = int(np.bitwise_xor.reduce((par.astype(np.int64) + 1) * (np.arange(par.size) + 7)))
    vec["kd_split_sum"] = int(sp.sum())
    r = 6.0
    counts, offsets, idx, key = t.range_batch(r, qs)
    vec["range_r"] = r
    vec["range_counts"] = counts.tolist()
    vec["range_idx_sorted_q0"] = sorted(idx[offsets[0]:offsets[1]].tolist())
    vec["range_key_sum_hex"] = float(np.sort(key).sum()).hex()
    ni, nd = t.nearest_batch(qs)
    vec["nearest_idx"] = ni.tolist()
    vec["nearest_dist_hex"] = [float(x).hex() for x in nd]
    # segment vs sphere on building2
    sph, ns = oracle.make_spheres(c, rad)
    segs = W.uniform_points(12, 400, [-20] * 3, [20] * 3).reshape(200, 2, 3)
    flags = []
    for s in segs:
        flags.append(int(L.orc_edge_check_all(sph, ns, 0, oracle._p(np.ascontiguousarray(s[0]), oracle.c_f64p),
                                              oracle._p(np.ascontiguousarray(s[1]), oracle.c_f64p), 0.5, 0)))
    vec["edge_flags_building2"] = flags
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)
    print("wrote fixtures:", sum(flags), "colliding segments of", len(flags))


if __name__ == "__main__":
    main()
