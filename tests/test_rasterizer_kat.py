"""Known-answer test of the analytic rasterizer: the CELLS the device's rasterizer code accumulates for a
set of edges equal the cells of the reference's AnalyticRasterizer (raster/analyticrasterizer_p.h:289-1210, instantiated
unmodified by oracle/ref_internals.cpp).  Both ways a GPU lane can reach a scanline are checked - stepping through the
rows of an edge, and jumping to a row with edge_advance_to_y (the unit of work of a lane in k_tile_render) - which is the
property the reference's own unit test pins (raster/analyticrasterizer_test.cpp:34-157: advanceToY == stepping).
The device code is dev_raster.cuh compiled for the host by tests/hostsim; the GPU build of the same header is covered by
the pixel parity tests, and a committed digest keeps this test meaningful where the reference is not built."""
import hashlib

import numpy as np
import pytest

from tests import hostsim

W = H = 1000


def random_lines(seed, n):
    """Like the reference's unit test: endpoints uniform over the raster in 24.8 fixed point, plus the shapes that take the
    special paths of prepare(): vertical, single-cell, nearly horizontal, horizontal (skipped) and right-to-left lines."""
    rng = np.random.default_rng(seed)
    ln = (rng.random((n, 4)) * [W * 256, H * 256, W * 256, H * 256]).astype(np.int32)
    k = n // 10
    ln[0:k, 2] = ln[0:k, 0]                                               # vertical
    ln[k:2 * k, 2] = ln[k:2 * k, 0] + rng.integers(-255, 256, k)          # steep, within a cell or two
    ln[2 * k:3 * k, 3] = ln[2 * k:3 * k, 1] + rng.integers(-300, 301, k)  # nearly horizontal (many cells per scanline)
    ln[3 * k:3 * k + 8, 3] = ln[3 * k:3 * k + 8, 1]                       # horizontal: contributes nothing
    ln[3 * k + 8:4 * k, :] &= ~0xFF                                       # endpoints on pixel corners
    return np.clip(ln, 0, [W * 256, H * 256, W * 256, H * 256]).astype(np.int32)


@pytest.fixture(scope="module")
def refint(ref):
    from oracle import ref_internals as RI
    if not RI.available():
        pytest.skip("oracle/_ref/libref_internals.so not built")
    return RI


def device_cells(lines, mode):
    cells = np.zeros((H, W + 2), np.uint32)
    ln = np.ascontiguousarray(lines, np.int32)
    hostsim.lib().hostsim_rasterize_cells(ln.ctypes.data, len(ln), W, H, mode, cells.ctypes.data)
    return cells


@pytest.mark.parametrize("seed,n", [(0x1234, 20000), (77, 5000)])
def test_cells_equal_reference_rasterizer(refint, seed, n):
    lines = random_lines(seed, n)
    want = refint.rasterize_edges(lines, W, H)
    assert want.any()
    assert np.array_equal(device_cells(lines, 0), want), "stepping through the rows differs from AnalyticRasterizer"
    assert np.array_equal(device_cells(lines, 1), want), "jumping to a row (edge_advance_to_y) differs from AnalyticRasterizer"


def test_jumping_equals_stepping_and_digest():
    """Runs everywhere (no reference needed): advanceToY == stepping on 50 000 edges, and the cells' digest is the one the
    reference produced when this test was written."""
    lines = random_lines(0x1234, 50000)
    a, b = device_cells(lines, 0), device_cells(lines, 1)
    assert np.array_equal(a, b)
    # every scanline's cells of a closed set of covers: the sum over a row equals the sum of (cover << 9) of its edges
    ln = lines.astype(np.int64)
    y0, y1 = np.minimum(ln[:, 1], ln[:, 3]), np.maximum(ln[:, 1], ln[:, 3])
    sign = np.where(ln[:, 1] > ln[:, 3], -1, 1)
    rows = np.arange(H, dtype=np.int64)[:, None] * 256
    cover = np.clip(np.minimum(y1[None, :], rows + 256) - np.maximum(y0[None, :], rows), 0, None) * sign[None, :]
    assert np.array_equal(a.sum(axis=1, dtype=np.uint64) & 0xFFFFFFFF, (cover.sum(axis=1) << 9) & 0xFFFFFFFF)
    assert hashlib.sha256(a.tobytes()).hexdigest() == DIGEST_50000


def test_err_multi_step_equals_single_steps():
    """The closed form behind edge_advance_to_y and the windowed one-row stepper: jumping `count` error steps at once
    (AnalyticUtils::acc_err_multi_step, analyticrasterizer_p.h:84-100) == stepping `count` times, 2 M random states."""
    assert hostsim.lib().hostsim_err_multi_step_check(12345, 2_000_000) == 0


DIGEST_50000 = "a8e2b642fea5d2ee3fa61c615874990882e1d95b54d92459e9b7fbb5aff123aa"
