"""CPU parity: the host simulator (tests/hostsim - the kernels' __host__ __device__ headers executed in the kernels'
per-tile order) against the UNMODIFIED reference, on wider scenes than the golden set.  This is what lets kernel logic
be debugged in this GPU-less container; the GPU suite repeats the comparison with the real kernels."""
import numpy as np
import pytest

from tests import scenes as S

W0, H0 = 512, 600


def compare(ref, scene, W, H, fmt=1, seed=1, max_diff=0):
    from tests import hostsim
    ri, _ = S.draw(ref, scene, W, H, fmt, seed)
    n, d = S.channel_diff(ri.to_numpy(), hostsim.draw(scene, W, H, fmt, seed))
    assert d <= max_diff, f"{n} pixels differ, max channel diff {d}"
    if max_diff == 0:
        assert n == 0


@pytest.mark.parametrize("kind", ["A", "U"])
def test_bl_bench_rects(ref, kind):
    compare(ref, S.rects(kind, 300, 64, W0, H0), W0, H0)


@pytest.mark.parametrize("rule", [0, 1])
def test_polygons(ref, rule):
    compare(ref, S.polygons(200, 128, 20, W0, H0, rule), W0, H0)


@pytest.mark.parametrize("size", [(700, 300), (1300, 200), (513, 97), (2000, 64)])
def test_shallow_slivers(ref, size):
    """Nearly horizontal edges: the window skip of the one-row stepper (edge_step_scanline<kWindow>)."""
    compare(ref, S.slivers(150, *size), *size)


@pytest.mark.parametrize("kind", ["quad", "cubic"])
def test_curves_clipped_by_the_canvas(ref, kind):
    compare(ref, S.curve_paths(kind, 150, W0, H0, 1), W0, H0)


@pytest.mark.parametrize("style,tol", [("linear", 0), ("radial", 0), ("conic", 1)])
def test_gradients(ref, style, tol):
    compare(ref, S.polygons(80, 200, 10, W0, H0, 0, style, 2), W0, H0, max_diff=tol)


@pytest.mark.parametrize("quality", [0, 1])
def test_patterns(ref, quality):
    compare(ref, S.pattern_shapes("rot", 80, 96, W0, H0, quality, 1), W0, H0)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_mixed_unaligned_canvas(ref, fmt):
    compare(ref, S.mixed(200, 513, 257), 513, 257, fmt, 9, max_diff=1)


@pytest.mark.parametrize("style", ["solid", "radial"])
@pytest.mark.parametrize("fmt", [1, 3])
def test_fill_mask(ref, style, fmt):
    compare(ref, S.masked_fills(80, 300, 200, S.SRC_OVER, style), 300, 200, fmt, 4)


@pytest.mark.parametrize("kind", [3, 4])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_node_expansion_equals_sequential_flattening(kind, mode):
    """k_build_edges flattens a curve with a warp: the monotone pieces are expanded breadth first into <= 32 nodes of the
    reference's subdivision tree and every node is walked on its own (dev_flatten.cuh node_split / node_walk).  The
    same device functions, run on the CPU, must emit the edges of the sequential walk (edgebuilder_p.h:2029-2445): equal
    multisets of lines, equal signed coverage of the (un-merged) vertical border lines.  mode: 0 control points slightly
    outside the clip box, 1 far outside, 2 all inside, 3 glyph-sized curves."""
    from tests import hostsim
    assert hostsim.flatten_equivalence(kind, 60000, 99 + mode, mode) == 0
