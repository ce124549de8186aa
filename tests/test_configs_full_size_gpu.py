"""BASELINE.json configs 0, 2 and 3 at their full size, end to end through the drop-in: ONE Blend2D application
(shim/bl_scene_driver.cpp, public bl_* calls only) draws the scene on a GPU context of shim/_build/libblend2d_gpu.so and
on the unmodified reference (oracle/_ref), and the two images must be identical.  These are the parity gates bench.py
applies before it times a configuration, as tests (config 1: tests/test_full_size_gpu.py and the bench gate; config 4:
tests/test_sharding.py, test_full_size_gpu.py and the frames leg's gate)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W4K, H4K = 3840, 2160


@pytest.fixture(scope="module")
def gb(ref):
    import bench
    if bench.Driver.load(bench.REF_DRIVER) is None:
        pytest.skip("oracle/_ref/libref_scene_driver.so not built")
    return bench.GpuBench(types.SimpleNamespace(no_parity=False, queue_limit=0))


def both(gb, scene, n, W, H):
    import bench
    sess = gb.open(scene, W, H, queue_limit=0)
    sess.clear(); sess.draw(0, n); sess.flush(True)
    assert sess.error_flags() == 0
    gpu = sess.pixels().copy()
    sess.close()
    ref = bench.run_reference(scene, n, W, H, 1, 0, bench.host_threads(), return_pixels=True)
    assert ref is not None
    return ref["pixels"], gpu


def assert_identical(cpu, gpu):
    import bench
    n, d = bench.channel_diff(cpu, gpu)
    assert (n, d) == (0, 0), f"{n} pixels differ, max channel diff {d}"
    assert gpu.any(), "nothing was drawn"


def test_config0_bl_bench_rects_20000(gb):
    import bench_scenes as BS
    scene, keep = BS.make_config0_scene(20000)
    assert_identical(*both(gb, scene, 20000, 512, 600))


def test_config2_patterns_10000(gb):
    """The reference's portable pipeline has no Plus / Multiply / Screen: same geometry, sprites and fetchers with those
    operators remapped (bench.py does the same); the operators themselves: tests/test_oracle.py."""
    import bench_scenes as BS
    scene, keep = BS.make_config2_scene(10000, W4K, H4K)
    pscene, pkeep = BS.slice_fills(scene, keep, {BS.PLUS: BS.SRC_OVER, BS.MULTIPLY: BS.SRC_COPY, BS.SCREEN: BS.SRC_OVER})
    assert_identical(*both(gb, pscene, 10000, W4K, H4K))


def test_config3_100000_glyphs(gb):
    import bench_scenes as BS
    scene, keep = BS.make_config3_scene(25000, W4K, H4K)
    assert_identical(*both(gb, scene, 25000, W4K, H4K))
