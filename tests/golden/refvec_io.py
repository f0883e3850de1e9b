"""Text container shared by make_reference_inputs.py, julia/make_reference_vectors.jl and
tests/test_reference_vectors.py: named 2-D arrays, floats as the 16 hex digits of their IEEE-754 bit pattern (no
decimal round trip anywhere), integers in decimal.

    @name rows cols f|i
    <rows lines of cols tokens>
"""
import numpy as np


def write_arrays(path, arrays):
    with open(path, "w") as f:
        for name, a in arrays.items():
            a = np.asarray(a)
            if a.ndim == 1:
                a = a.reshape(-1, 1) if a.size else a.reshape(0, 1)
            kind = "f" if a.dtype.kind == "f" else "i"
            f.write(f"@{name} {a.shape[0]} {a.shape[1]} {kind}\n")
            if kind == "f":
                bits = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
                for row in bits:
                    f.write(" ".join(f"{int(v):016x}" for v in row) + "\n")
            else:
                for row in a.astype(np.int64):
                    f.write(" ".join(str(int(v)) for v in row) + "\n")


def read_arrays(path):
    out = {}
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        i += 1
        if not ln.startswith("@"):
            continue
        name, rows, cols, kind = ln[1:].split()
        rows, cols = int(rows), int(cols)
        if kind == "f":
            a = np.zeros((rows, cols), dtype=np.uint64)
            for r in range(rows):
                a[r] = [int(t, 16) for t in lines[i + r].split()]
            out[name] = a.view(np.float64)
        else:
            a = np.zeros((rows, cols), dtype=np.int64)
            for r in range(rows):
                a[r] = [int(t) for t in lines[i + r].split()]
            out[name] = a
        i += rows
    return out
