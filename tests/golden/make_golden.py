"""Regenerates tests/golden/*.npz from the reference's data fixture (run in the build container only).

/root/reference does not exist on the GPU box, so the MovieLens-100K fixture the reference's own
tests run on (core/base_test.go:50-64 via core/data.go:270 LoadDataFromBuiltIn("ml-100k")) is stored
here in compressed form: u.data (100,000 rows, md5 6e47046882bad158b0efbb84cd5cb987) and the five
fixed folds u1..u5 (.base/.test) that ship beside it in core/data/ml-100k/.  Parsing follows
core/data.go:298-309: tab-separated, fields 0..2 through Atoi.
MovieLens data: GroupLens Research, University of Minnesota (see core/data/ml-100k/README).
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/core/data/ml-100k")
OUT = Path(__file__).resolve().parent


def load(name):
    rows = [ln.split("\t") for ln in (SRC / name).read_text().splitlines()]
    a = np.array([[int(r[0]), int(r[1]), int(r[2])] for r in rows], dtype=np.int32)
    return a


if __name__ == "__main__":
    md5 = hashlib.md5((SRC / "u.data").read_bytes()).hexdigest()
    assert md5 == "6e47046882bad158b0efbb84cd5cb987", md5
    arrays = {"u_data": load("u.data")}
    for f in range(1, 6):
        arrays[f"u{f}_base"] = load(f"u{f}.base")
        arrays[f"u{f}_test"] = load(f"u{f}.test")
    # u16 columns compress better; max id 1682, ratings 1..5
    packed = {k: v.astype(np.uint16) for k, v in arrays.items()}
    np.savez_compressed(OUT / "ml100k.npz", **packed)
    print({k: v.shape for k, v in arrays.items()})
