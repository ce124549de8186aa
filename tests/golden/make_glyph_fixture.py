"""Generates tests/golden/glyph_run.npz - config 3 (FillGlyphRun): text rendered by the UNMODIFIED reference with its
bundled test font, plus the glyph-run outlines the reference itself decodes for each string.

    python -m tests.golden.make_glyph_fixture          (needs /root/reference and oracle/_ref)

Scene = the reference tester's text workload (blend2d-testing/tests/bl_test_context_utilities.h:1160-1222): 4-character
strings from an 84-character alphabet, font size 20, ABeeZee-Regular.  For every string the fixture stores
  * the path returned by bl_font_get_glyph_run_outlines() with the string's origin as user transform - exactly the
    geometry fill_glyph_run_sink (raster/rastercontextops.cpp:51-59) hands to the edge builder, and
  * nothing else of the font: glyph decoding stays on the host.
`expected` is what bl_context_fill_utf8_text_d() rendered.  Filling the stored outlines with fill_path() reproduces it
bit for bit on the reference (checked here), so the GPU path can be compared with real text output.
"""
import ctypes as C
import os

import numpy as np

from oracle import ref_blend2d as R

W, H, STRINGS, SIZE = 640, 360, 200, 20.0
ALPHABET = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789!@#$%^&*()[]{}<>?/+-=_.,"
FONT = b"/root/reference/blend2d-testing/resources/ABeeZee-Regular.ttf"


class Mat(C.Structure):
    _fields_ = [("m", C.c_double * 6)]


def main():
    L = R.lib()
    Core, P = R.Core, C.POINTER
    L.bl_font_face_init.argtypes = [P(Core)]
    L.bl_font_face_create_from_file.argtypes = [P(Core), C.c_char_p, C.c_uint32]
    L.bl_font_init.argtypes = [P(Core)]
    L.bl_font_create_from_face.argtypes = [P(Core), P(Core), C.c_float]
    L.bl_glyph_buffer_init.argtypes = [P(Core)]
    L.bl_glyph_buffer_set_text.argtypes = [P(Core), C.c_char_p, C.c_size_t, C.c_uint32]
    L.bl_font_shape.argtypes = [P(Core), P(Core)]
    L.bl_glyph_buffer_get_glyph_run.restype = C.c_void_p
    L.bl_glyph_buffer_get_glyph_run.argtypes = [P(Core)]
    L.bl_font_get_glyph_run_outlines.argtypes = [P(Core), C.c_void_p, C.c_void_p, P(Core), C.c_void_p, C.c_void_p]
    L.bl_path_get_size.restype = C.c_size_t
    L.bl_path_get_size.argtypes = [P(Core)]
    L.bl_path_get_command_data.restype = P(C.c_uint8)
    L.bl_path_get_command_data.argtypes = [P(Core)]
    L.bl_path_get_vertex_data.restype = P(C.c_double)
    L.bl_path_get_vertex_data.argtypes = [P(Core)]
    L.bl_context_fill_utf8_text_d.argtypes = [P(Core), P(R.Point), P(Core), C.c_char_p, C.c_size_t]

    face, font, gb = Core(), Core(), Core()
    L.bl_font_face_init(C.byref(face))
    assert L.bl_font_face_create_from_file(C.byref(face), FONT, 0) == 0
    L.bl_font_init(C.byref(font))
    assert L.bl_font_create_from_face(C.byref(font), C.byref(face), SIZE) == 0
    L.bl_glyph_buffer_init(C.byref(gb))

    rng = np.random.default_rng(77)
    img = R.Image(W, H, 1)
    ctx = R.Context(img)
    check = R.Image(W, H, 1)
    cctx = R.Context(check)
    cmds_all, vtx_all, offsets, colors = [], [], [0], []
    for i in range(STRINGS):
        text = bytes(ALPHABET[int(k)] for k in rng.integers(0, len(ALPHABET), 4))
        x, y = float(rng.uniform(-10, W - 30)), float(rng.uniform(5, H + 5))
        color = int(rng.integers(0, 2 ** 32)) | 0x40000000
        ctx.set_fill_style(color)
        pt = R.Point(x, y)
        assert L.bl_context_fill_utf8_text_d(C.byref(ctx._c), C.byref(pt), C.byref(font), text, len(text)) == 0
        # the outlines the glyph-run sink sees
        assert L.bl_glyph_buffer_set_text(C.byref(gb), text, len(text), 0) == 0
        assert L.bl_font_shape(C.byref(font), C.byref(gb)) == 0
        run = L.bl_glyph_buffer_get_glyph_run(C.byref(gb))
        p = R.Path()
        m = Mat((C.c_double * 6)(1, 0, 0, 1, x, y))
        assert L.bl_font_get_glyph_run_outlines(C.byref(font), run, C.byref(m), C.byref(p._c), None, None) == 0
        n = L.bl_path_get_size(C.byref(p._c))
        cmds_all.append(np.ctypeslib.as_array(L.bl_path_get_command_data(C.byref(p._c)), (n,)).copy())
        vtx_all.append(np.ctypeslib.as_array(L.bl_path_get_vertex_data(C.byref(p._c)), (n * 2,)).copy())
        offsets.append(offsets[-1] + n)
        colors.append(color)
        cctx.set_fill_style(color)
        cctx.fill_path(p)
    ctx.end(); cctx.end()
    expected = img.to_numpy().copy()
    assert np.array_equal(expected, check.to_numpy()), "fill_path(outlines) must equal fill_utf8_text on the reference"
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glyph_run.npz")
    np.savez_compressed(path, cmds=np.concatenate(cmds_all), vtx=np.concatenate(vtx_all), offsets=np.array(offsets, dtype=np.int64),
                        colors=np.array(colors, dtype=np.uint32), expected=expected, size=np.array([W, H]))
    print(f"wrote {path}: {STRINGS} strings, {offsets[-1]} vertices, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
