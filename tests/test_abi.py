"""The drop-in boundary: blend2d_b200/libb2dgpu.so must load without a GPU and export every function that include/*.h
declares; the plain-C structs must have the sizes the headers document (they mirror reference structs whose sizes were
probed on x86-64).  No compute entry point is called here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", h) for h in ("b2dgpu.h", "b2d_host.h")]
SO = os.path.join(ROOT, "blend2d_b200", "libb2dgpu.so")


def declared_functions():
    names = []
    for h in HEADERS:
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"B2DGPU_API\s+[\w\s\*]+?\b(\w+)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_the_boundary():
    names = declared_functions()
    for must in ("b2dgpu_runtime_create", "b2dgpu_runtime_get", "b2dgpu_runtime_test", "b2dgpu_submit", "b2dgpu_sync",
                 "b2dgpu_runtime_destroy", "b2d_context_create", "b2d_context_fill_path_d", "b2d_scene_replay"):
        assert must in names
    assert len(names) >= 50


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(SO)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_python_binding_table_matches_headers():
    from blend2d_b200 import _native as N
    assert sorted(N.EXPORTED_SYMBOLS) == declared_functions()


def test_no_unexpected_exports():
    out = subprocess.check_output(["nm", "-D", "--defined-only", SO], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    ours = {n for n in exported if n.startswith(("b2d_", "b2dgpu_"))}
    assert ours == set(declared_functions())


def test_struct_sizes_match_the_headers(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "b2dgpu.h"
#include "b2d_host.h"
#include "b2d_scene.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b2dgpu_fetch_data), sizeof(b2dgpu_command), sizeof(b2dgpu_edge),
         sizeof(b2dgpu_segment), sizeof(b2dgpu_geometry_state), sizeof(b2dgpu_dispatch_data),
         offsetof(b2dgpu_fetch_data, pattern.simple), offsetof(b2dgpu_fetch_data, gradient.linear), sizeof(b2d_scene_fill));
  return 0;
}''')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    # FetchData 176 (Pattern.simple @32, Gradient.linear @16), RenderCommand-sized command 64, EdgePoint pair 16,
    # DispatchData 16 - SURVEY.md section 8 "ABI sizes [probe, x86-64]"
    assert got == [176, 64, 16, 12, 96, 16, 32, 16, 160]


REFERENCE = "/root/reference"

LAYOUT_CHECK = r'''
// Every field the device reads, at the offset the reference's own headers give it (compile-time only).
#include <blend2d/core/api-build_p.h>
#include <blend2d/pipeline/pipedefs_p.h>
#include <blend2d/raster/edgestorage_p.h>
#include <stddef.h>
#include "b2dgpu.h"
using FD = bl::Pipeline::FetchData;
#define SAME_OFF(ours, theirs) static_assert(offsetof(b2dgpu_fetch_data, ours) == offsetof(FD, theirs), #ours " vs " #theirs)
static_assert(sizeof(b2dgpu_fetch_data) == sizeof(FD), "FetchData size");
static_assert(sizeof(b2dgpu_dispatch_data) == sizeof(bl::Pipeline::DispatchData), "DispatchData size");
static_assert(sizeof(b2dgpu_edge) == 2 * sizeof(bl::RasterEngine::EdgePoint<int>), "an edge is two EdgePoint<int>");
SAME_OFF(solid.prgb32, solid.prgb32);
SAME_OFF(pattern.src.pixel_data, pattern.src.pixel_data); SAME_OFF(pattern.src.stride, pattern.src.stride); SAME_OFF(pattern.src.w, pattern.src.size);
SAME_OFF(pattern.simple.tx, pattern.simple.tx); SAME_OFF(pattern.simple.ty, pattern.simple.ty);
SAME_OFF(pattern.simple.rx, pattern.simple.rx); SAME_OFF(pattern.simple.ry, pattern.simple.ry);
SAME_OFF(pattern.simple.wa, pattern.simple.wa); SAME_OFF(pattern.simple.wd, pattern.simple.wd);
SAME_OFF(pattern.affine.xx, pattern.affine.xx); SAME_OFF(pattern.affine.yy, pattern.affine.yy); SAME_OFF(pattern.affine.tx, pattern.affine.tx);
SAME_OFF(pattern.affine.ox, pattern.affine.ox); SAME_OFF(pattern.affine.rx, pattern.affine.rx); SAME_OFF(pattern.affine.xx2, pattern.affine.xx2);
SAME_OFF(pattern.affine.min_x, pattern.affine.min_x); SAME_OFF(pattern.affine.max_y, pattern.affine.max_y);
SAME_OFF(pattern.affine.cor_x, pattern.affine.cor_x); SAME_OFF(pattern.affine.tw, pattern.affine.tw); SAME_OFF(pattern.affine.addr_mul32, pattern.affine.addr_mul32);
SAME_OFF(gradient.lut.data, gradient.lut.data); SAME_OFF(gradient.lut.size, gradient.lut.size);
SAME_OFF(gradient.linear.pt, gradient.linear.pt); SAME_OFF(gradient.linear.dy, gradient.linear.dy); SAME_OFF(gradient.linear.dt, gradient.linear.dt);
SAME_OFF(gradient.linear.maxi, gradient.linear.maxi); SAME_OFF(gradient.linear.rori, gradient.linear.rori);
SAME_OFF(gradient.radial.tx, gradient.radial.tx); SAME_OFF(gradient.radial.yx, gradient.radial.yx); SAME_OFF(gradient.radial.amul4, gradient.radial.amul4);
SAME_OFF(gradient.radial.inv2a, gradient.radial.inv2a); SAME_OFF(gradient.radial.sq_fr, gradient.radial.sq_fr); SAME_OFF(gradient.radial.sq_inv2a, gradient.radial.sq_inv2a);
SAME_OFF(gradient.radial.b0, gradient.radial.b0); SAME_OFF(gradient.radial.dd0, gradient.radial.dd0); SAME_OFF(gradient.radial.by, gradient.radial.by);
SAME_OFF(gradient.radial.ddy, gradient.radial.ddy); SAME_OFF(gradient.radial.f32_ddd, gradient.radial.f32_ddd); SAME_OFF(gradient.radial.f32_bd, gradient.radial.f32_bd);
SAME_OFF(gradient.radial.maxi, gradient.radial.maxi); SAME_OFF(gradient.radial.rori, gradient.radial.rori);
SAME_OFF(gradient.conic.tx, gradient.conic.tx); SAME_OFF(gradient.conic.yx, gradient.conic.yx); SAME_OFF(gradient.conic.q_coeff, gradient.conic.q_coeff);
SAME_OFF(gradient.conic.n_div_1_2_4, gradient.conic.n_div_1_2_4); SAME_OFF(gradient.conic.offset, gradient.conic.offset); SAME_OFF(gradient.conic.xx, gradient.conic.xx);
SAME_OFF(gradient.conic.maxi, gradient.conic.maxi); SAME_OFF(gradient.conic.rori, gradient.conic.rori);
// BLPipeSignature bit fields (pipedefs_p.h:232-366) behind the B2DGPU_SIG_* accessors
static_assert(B2DGPU_SIG_COMP_OP(uint32_t(bl::Pipeline::Signature::from_comp_op(bl::CompOpExt(17)).value)) == 17u, "comp op field");
static_assert(B2DGPU_SIG_FILL_TYPE(uint32_t(bl::Pipeline::Signature::from_fill_type(bl::Pipeline::FillType::kAnalytic).value)) == B2DGPU_FILL_ANALYTIC, "fill type field");
static_assert(B2DGPU_SIG_FETCH_TYPE(uint32_t(bl::Pipeline::Signature::from_fetch_type(bl::Pipeline::FetchType::kGradientConicNN).value)) == B2DGPU_FETCH_GRADIENT_CONIC_NN, "fetch type field");
static_assert(B2DGPU_SIG_DST_FORMAT(uint32_t(bl::Pipeline::Signature::from_dst_format(bl::FormatExt::kPRGB32).value)) == B2DGPU_FORMAT_PRGB32, "dst format field");
static_assert(B2DGPU_SIG_SRC_FORMAT(uint32_t(bl::Pipeline::Signature::from_src_format(bl::FormatExt::kA8).value)) == B2DGPU_FORMAT_A8, "src format field");
int main() { return 0; }
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "blend2d")), reason="needs the reference headers (/root/reference)")
def test_struct_layouts_match_the_reference_headers(tmp_path):
    """include/b2dgpu.h mirrors FetchData / DispatchData / EdgePoint / BLPipeSignature: checked field by field against the
    reference's private headers themselves (they compile standalone), not against remembered numbers."""
    src = tmp_path / "layout.cpp"
    src.write_text(LAYOUT_CHECK)
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-w", "-DBL_STATIC", "-DBL_BUILD_NO_JIT", "-I", REFERENCE,
                           "-I", os.path.join(ROOT, "include"), str(src)])


def test_signature_queries_work_without_a_device():
    """PipeRuntime::test/get semantics are host logic: NOT_IMPLEMENTED / NO_ENTRY for signatures outside the table."""
    from blend2d_b200 import _native as N
    assert N.lib.b2dgpu_abi_version() >= 1


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: creating a runtime on a machine without a GPU must return an error, not render on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import blend2d_b200 as G
    img = G.Image(16, 16, 1)
    with pytest.raises(Exception):
        G.Context(img)


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "blend2d_b200")
    pat = re.compile(r"(from|import)\s+(oracle|tests)\b|#include\s+\"[^\"]*(oracle|hostsim)")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), f"{f} reaches into test infrastructure"
