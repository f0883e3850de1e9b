#!/usr/bin/env python
"""Collects the numbers the REFERENCE ITSELF left in its tree into tests/golden/reference_bvp_dists.json.

R/BVPData/BVPDistFromStaticObs_<agent>.txt were written by the reference's saveBVPDists (R/DRRT_Q.jl:91-97) from
findClosestObs / findClosestObsToLineSegment (R/DRRT_Q.jl:127-160): for every recorded point p
    [p, min_j sqrt((p1-c_j1)^2 + (p2-c_j2)^2 + (p3-c_j3)^2) - radius_(argmin)]      (first minimum wins)
over the static sphere obstacles of the run.  The obstacle file of that run is not recorded; this script tries every
sphere-format file under R/environments and keeps the one that reproduces ALL rows bit for bit with IEEE double
arithmetic (R/environments/buildingsSmall.txt, 15 spheres) -- so the rows are a known-answer test, produced by Julia,
of the reference's Euclidean distance arithmetic (operation order, correctly rounded sqrt) that euclidianDist
(R/DRRT_distance_functions.jl) shares and every kd key, range decision and point-check certificate is made of.

usage: python tests/golden/make_reference_outputs.py /root/reference/code_RRTQx_3D
"""
import glob
import json
import math
import os
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/code_RRTQx_3D"


def parse_spheres(path):
    lines = [l.strip() for l in open(path) if l.strip()]
    try:
        n = int(lines[0])
        out, i = [], 1
        for _ in range(n):
            c = [float(t) for t in lines[i].split(",")]
            if len(c) != 3:
                return None
            out.append(c + [float(lines[i + 1])])
            int(lines[i + 2])
            i += 3
        return out
    except Exception:
        return None


rows = []
for f in sorted(glob.glob(os.path.join(R, "BVPData", "BVPDistFromStaticObs_*.txt"))):
    for line in open(f):
        if line.strip():
            rows.append({"file": os.path.basename(f), "v": [float(t) for t in line.split(",")]})


def closest(p, spheres):
    best, rad, arg = -1.0, 0.0, -1
    for j, o in enumerate(spheres):
        d = math.sqrt((p[0] - o[0]) ** 2 + (p[1] - o[1]) ** 2 + (p[2] - o[2]) ** 2)
        if d < best or best == -1.0:
            best, rad, arg = d, o[3], j
    return best - rad, arg


match = None
for env in sorted(glob.glob(os.path.join(R, "environments", "*.txt"))):
    S = parse_spheres(env)
    if S and all(closest(r["v"][:3], S)[0] == r["v"][3] for r in rows):
        match = (env, S)
        break
assert match, "no environment file reproduces the recorded distances"
env, S = match
out = {"source": "R/BVPData/BVPDistFromStaticObs_*.txt (reference output), obstacles R/environments/" + os.path.basename(env),
       "spheres_hex": [[x.hex() for x in s] for s in S],
       "rows_hex": [{"file": r["file"], "v": [x.hex() for x in r["v"]], "argmin": closest(r["v"][:3], S)[1]} for r in rows]}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_bvp_dists.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, len(S), "spheres", len(rows), "rows from", os.path.basename(env))
