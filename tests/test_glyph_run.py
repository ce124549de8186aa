"""Config 3 (FillGlyphRun): text rendered by the unmodified reference (bl_context_fill_utf8_text_d, bundled ABeeZee font)
against our path filling the glyph-run outlines the reference decoded for the same strings
(tests/golden/make_glyph_fixture.py).  Glyph decoding is host work outside the hot path; everything
from the outline on - flattening of the quadratic TrueType curves, coverage, compositing - is the path under test."""
import os

import numpy as np
import pytest

F = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glyph_run.npz"))
W, H = (int(v) for v in F["size"])


def build_path(api, cmds, vtx):
    """BLPathCmd: 0 move, 1 on-point (line / curve end), 2 quad control, 3 conic, 4 cubic control, 5 close."""
    p = api.Path()
    i, n = 0, len(cmds)
    while i < n:
        c = int(cmds[i])
        if c == 0:
            p.move_to(vtx[2 * i], vtx[2 * i + 1]); i += 1
        elif c == 1:
            p.line_to(vtx[2 * i], vtx[2 * i + 1]); i += 1
        elif c == 2:
            p.quad_to(vtx[2 * i], vtx[2 * i + 1], vtx[2 * i + 2], vtx[2 * i + 3]); i += 2
        elif c == 4:
            p.cubic_to(vtx[2 * i], vtx[2 * i + 1], vtx[2 * i + 2], vtx[2 * i + 3], vtx[2 * i + 4], vtx[2 * i + 5]); i += 3
        elif c == 5:
            p.close(); i += 1
        else:
            raise AssertionError(f"unexpected path command {c}")
    return p


def text_scene(api, ctx, rng=None):
    off = F["offsets"]
    for k in range(len(off) - 1):
        a, b = int(off[k]), int(off[k + 1])
        ctx.set_fill_style(int(F["colors"][k]))
        ctx.fill_path(build_path(api, F["cmds"][a:b], F["vtx"][2 * a:2 * b]))


def test_fixture_has_curves():
    assert (F["cmds"] == 2).sum() > 1000 and (F["expected"] != 0).sum() > 5000


def test_hostsim_renders_the_reference_text():
    from tests import hostsim
    got = hostsim.draw(text_scene, W, H, 1, 0)
    assert np.array_equal(got, F["expected"])


def test_reference_fill_path_equals_its_text(ref):
    img = ref.Image(W, H, 1)
    ctx = ref.Context(img)
    text_scene(ref, ctx)
    ctx.end()
    assert np.array_equal(img.to_numpy(), F["expected"])


@pytest.mark.gpu
def test_gpu_renders_the_reference_text(gpu):
    img = gpu.Image(W, H, 1)
    ctx = gpu.Context(img)
    text_scene(gpu, ctx)
    ctx.end()
    got = img.to_numpy().copy()
    ctx.close()
    assert np.array_equal(got, F["expected"])
