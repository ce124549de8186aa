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
