"""Generates tests/golden/golden.npz by running the UNMODIFIED reference (oracle/_ref/libblend2d_ref.so, built from
/root/reference by oracle/Makefile.ref) on the seeded cases of tests/golden/cases.py.

    python -m tests.golden.make_golden

The reference ships no golden images of its own (its tests compare two pipelines of the same build),
so these files are "outputs of the reference itself run here".  They let the CPU suite and the GPU suite check parity
even where oracle/_ref is not available.
"""
import os

import numpy as np

from oracle import ref_blend2d as R
from tests import scenes as S
from tests.golden.cases import CASES, W, H


def main():
    out = {}
    for name, (factory, fmt, seed, _tol) in CASES.items():
        img, _ = S.draw(R, factory(), W, H, fmt, seed)
        out[name] = img.to_numpy().copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
