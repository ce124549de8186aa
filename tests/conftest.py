import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")
    # A fresh checkout has no built library yet: build it (nvcc cross-compiles sm_100a without a GPU) instead of failing
    # every test at import time.  __graft_entry__.build() does the same and more.
    if not os.path.exists(os.path.join(ROOT, "blend2d_b200", "libb2dgpu.so")):
        import subprocess
        try:
            subprocess.check_call(["make", "-s", "-C", ROOT])
        except (OSError, subprocess.CalledProcessError) as e:
            # no nvcc here: the tests that load the library fail at their own import, the oracle / table tests still run
            sys.stderr.write("conftest: could not build libb2dgpu.so (%s)\n" % e)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference, built from /root/reference into oracle/_ref (travels to the GPU box)."""
    from oracle import ref_blend2d as R
    if not R.available() and os.path.isdir("/root/reference/blend2d"):
        # first use on a fresh checkout: build the unmodified reference with the committed recipe (about a minute)
        import subprocess
        subprocess.call(["make", "-s", "-j8", "-f", os.path.join(ROOT, "oracle", "Makefile.ref")], cwd=os.path.join(ROOT, "oracle"),
                        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if not R.available():
        pytest.skip("oracle/_ref/libblend2d_ref.so not built (needs /root/reference): run make -f oracle/Makefile.ref")
    return R


@pytest.fixture(scope="session")
def gpu():
    import blend2d_b200 as G
    return G
