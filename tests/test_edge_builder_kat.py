"""Intermediate known-answer test for K1: the edges OUR edge builder produces for a path against
the edges the reference's `EdgeBuilder<int>` produces for the same path, transform, clip box and tolerance
(oracle/ref_internals.cpp runs raster/edgebuilder_p.h:934-1070 of the unmodified reference).

Only the multiset of integer lines matters to the analytic rasterizer (it sums cover / area contributions), so the two
are compared in a canonical form: non-vertical lines as a sorted multiset in their original direction, vertical lines -
the reference merges consecutive clip-border intervals before truncating them, we emit them un-merged
(dev_flatten.cuh) - as the signed coverage along y they add up to per x.

CPU suite: the host instantiation of dev_flatten.cuh (tests/hostsim).  GPU suite: k_build_edges through
`b2dgpu_debug_build_edges`.
"""
import ctypes as C

import numpy as np
import pytest

from tests import scenes as S

W, H = 700, 500


def canonical(edges):
    e = np.asarray(edges, np.int64).reshape(-1, 4)
    e = e[e[:, 1] != e[:, 3]]                                        # horizontal lines carry no coverage
    vert = e[:, 0] == e[:, 2]
    lines = e[~vert]
    lines = lines[np.lexsort(lines.T[::-1])]
    runs = []
    v = e[vert]
    for x in np.unique(v[:, 0]):
        ev = {}
        for _, y0, _, y1 in v[v[:, 0] == x]:
            sgn = 1 if y0 < y1 else -1
            ev[min(y0, y1)] = ev.get(min(y0, y1), 0) + sgn
            ev[max(y0, y1)] = ev.get(max(y0, y1), 0) - sgn
        w, start = 0, 0
        for y in sorted(ev):
            if ev[y] == 0:
                continue
            if w != 0:
                runs.append((int(x), int(start), int(y), int(w)))
            w += ev[y]; start = y
        assert w == 0
    # merge adjacent runs of equal winding
    merged = []
    for r in runs:
        if merged and merged[-1][0] == r[0] and merged[-1][2] == r[1] and merged[-1][3] == r[3]:
            merged[-1] = (r[0], merged[-1][1], r[2], r[3])
        else:
            merged.append(r)
    return lines, merged


def random_paths(api, rng, n, margin):
    paths = []
    for i in range(n):
        p = api.Path()
        pt = lambda: (float(rng.uniform(-margin, W + margin)), float(rng.uniform(-margin, H + margin)))
        p.move_to(*pt())
        for _ in range(int(rng.integers(2, 9))):
            k = int(rng.integers(0, 4))
            if k == 0:
                p.line_to(*pt())
            elif k == 1:
                p.quad_to(*pt(), *pt())
            elif k == 3:
                p.conic_to(*pt(), *pt(), float(rng.choice([0.2, 0.7071067811865476, 1.0, 3.0, rng.uniform(0.05, 6.0)])))
            else:
                p.cubic_to(*pt(), *pt(), *pt())
        if i % 2:
            p.close()
        paths.append(p)
    return paths


def reference_edges(refint, view, path, ci):
    cmd = view.commands[ci]
    gs = view.geometry_states[cmd.state_index]
    pc, pv = path.arrays()
    return refint.build_edges(pv, pc, True, list(gs.m), gs.transform_type, list(gs.clip), gs.tolerance_sq, H)


def record(gpu, paths, transform):
    ctx = gpu.Context(gpu.Image(W, H, 1), record_only=True)
    ctx.set_fill_style(0xFFFFFFFF)
    if transform == "rotate":
        ctx.rotate(0.3, W / 2, H / 2)
    elif transform == "scale":
        ctx.scale(1.7, 0.6)
    for p in paths:
        ctx.fill_path(p)
    return ctx


def check(refint, view, paths, edges, begins):
    assert view.command_count == len(paths)
    for ci, p in enumerate(paths):
        ours = edges[begins[ci]:begins[ci + 1]]
        want = reference_edges(refint, view, p, ci)
        la, va = canonical(ours)
        lb, vb = canonical(want)
        assert la.shape == lb.shape and np.array_equal(la, lb), f"path {ci}: lines differ ({len(la)} vs {len(lb)})"
        assert va == vb, f"path {ci}: vertical border coverage differs"


@pytest.fixture(scope="module")
def refint(ref):
    from oracle import ref_internals as RI
    if not RI.available():
        pytest.skip("oracle/_ref/libref_internals.so not built")
    return RI


@pytest.mark.parametrize("transform", ["none", "rotate", "scale"])
@pytest.mark.parametrize("margin", [0.0, 40.0, 900.0])
def test_hostsim_edges_equal_reference_edge_builder(refint, transform, margin):
    import blend2d_b200 as G
    from tests import hostsim
    rng = np.random.default_rng(int(margin) + len(transform))
    paths = random_paths(G, rng, 40, margin)
    ctx = record(G, paths, transform)
    view = ctx.peek_batch()
    edges, begins = hostsim.build_edges(ctx)
    check(refint, view, paths, edges, begins)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("transform", ["none", "rotate", "scale"])
@pytest.mark.parametrize("margin", [0.0, 40.0, 900.0])
def test_gpu_edges_equal_reference_edge_builder(refint, gpu, transform, margin):
    from blend2d_b200 import _native as N
    rng = np.random.default_rng(7 + int(margin) + len(transform))
    paths = random_paths(gpu, rng, 120, margin)
    ctx = record(gpu, paths, transform)
    view = ctx.peek_batch()
    rt = gpu.Runtime(device=0)
    cap = 1 << 20
    buf = (N.Edge * cap)()
    count = C.c_uint32(0)
    begins = (C.c_uint32 * (view.command_count + 1))()
    N.check(N.lib.b2dgpu_debug_build_edges(rt._h, C.byref(view), buf, cap, C.byref(count), begins), "debug_build_edges")
    assert count.value <= cap
    edges = np.frombuffer(buf, dtype=np.int32).reshape(-1, 4)[:count.value].copy()
    check(refint, view, paths, edges, np.frombuffer(begins, dtype=np.uint32).copy())
    rt.close(); ctx.close()
