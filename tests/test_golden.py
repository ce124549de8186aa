"""Golden vectors: outputs of the unmodified reference on small seeded scenes (tests/golden/make_golden.py).

CPU suite: the reference build itself (when oracle/_ref exists) and the host simulator of the device functions
(tests/hostsim - the same __host__ __device__ headers the kernels are compiled from, executed in the kernels' per-tile
order) must reproduce them.  GPU suite: the CUDA path through the C-ABI must reproduce them.
"""
import os

import numpy as np
import pytest

from tests import scenes as S
from tests.golden.cases import CASES, W, H

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))


def check(name, got):
    n, d = S.channel_diff(GOLDEN[name], got)
    tol = CASES[name][3]
    assert d <= tol, f"{name}: {n} pixels differ, max channel diff {d}"
    if tol == 0:
        assert n == 0


def test_golden_file_is_complete():
    assert sorted(GOLDEN.files) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_reproduces_golden(ref, name):
    factory, fmt, seed, _ = CASES[name]
    img, _ = S.draw(ref, factory(), W, H, fmt, seed)
    assert np.array_equal(img.to_numpy(), GOLDEN[name])


@pytest.mark.parametrize("name", sorted(CASES))
def test_hostsim_reproduces_golden(name):
    from tests import hostsim
    factory, fmt, seed, _ = CASES[name]
    check(name, hostsim.draw(factory(), W, H, fmt, seed))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_golden(gpu, name):
    factory, fmt, seed, _ = CASES[name]
    img, ctx = S.draw(gpu, factory(), W, H, fmt, seed)
    got = img.to_numpy().copy()
    ctx.close()
    check(name, got)
