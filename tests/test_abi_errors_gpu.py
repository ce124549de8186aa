"""Error behaviour of the C-ABI on a device: BLResult codes instead of crashes for malformed batches, and a runtime
shared by two contexts driven from two threads (the reference's global pipeline runtime is shared the same way,
pipeline/jit/pipegenruntime.cpp:75-90)."""
import ctypes as C
import threading

import numpy as np
import pytest

from tests import scenes as S

pytestmark = pytest.mark.gpu

INVALID_VALUE, NOT_IMPLEMENTED = 0x10001, 0x10007      # BL_ERROR_INVALID_VALUE, BL_ERROR_NOT_IMPLEMENTED (core/api.h:1145-1152)


def simple_view(gpu, N, mutate):
    """A one-command batch (solid FillBoxA) with `mutate(cmd, view)` applied."""
    cmd = N.Command()
    cmd.type, cmd.signature, cmd.alpha = 1, 1 | (1 << 4) | (0 << 8) | (1 << 14), 255
    cmd.box[0], cmd.box[1], cmd.box[2], cmd.box[3] = 1, 1, 10, 10
    cmd.solid_prgb32 = 0xFF102030
    view = N.BatchView()
    view.struct_size = C.sizeof(N.BatchView)
    view.command_count = 1
    view.commands = C.pointer(cmd)
    mutate(cmd, view)
    return cmd, view


@pytest.mark.parametrize("name,mutate,code", [
    ("unknown command type", lambda c, v: setattr(c, "type", 77), INVALID_VALUE),
    ("alpha out of range", lambda c, v: setattr(c, "alpha", 300), INVALID_VALUE),
    ("empty box", lambda c, v: c.box.__setitem__(2, 1), INVALID_VALUE),
    ("fetch index out of range", lambda c, v: (setattr(c, "signature", c.signature | (15 << 16)), setattr(c, "fetch_index", 5)), INVALID_VALUE),
    ("operator outside the table (internal alpha inversion, 29)", lambda c, v: setattr(c, "signature", (c.signature & ~0x3F00) | (29 << 8)), NOT_IMPLEMENTED),
    ("edge range out of bounds", lambda c, v: (setattr(c, "type", 3), setattr(c, "signature", (c.signature & ~0xC000) | (3 << 14)), setattr(c, "data_count", 4)), INVALID_VALUE),
    ("bad struct size", lambda c, v: setattr(v, "struct_size", 8), INVALID_VALUE),
])
def test_malformed_batches_are_refused(gpu, name, mutate, code):
    from blend2d_b200 import _native as N
    img = gpu.Image(32, 32, 1)
    ctx = gpu.Context(img)
    cmd, view = simple_view(gpu, N, mutate)
    r = N.lib.b2dgpu_submit(ctx.runtime_handle(), ctx.target_handle(), C.byref(view))
    assert r == code, f"{name}: 0x{r:X} ({N.lib.b2dgpu_last_error_message().decode()})"
    ctx.end()
    assert not img.to_numpy().any()               # nothing was drawn
    ctx.close()


def test_null_arguments(gpu):
    from blend2d_b200 import _native as N
    img = gpu.Image(8, 8, 1)
    ctx = gpu.Context(img)
    assert N.lib.b2dgpu_submit(None, ctx.target_handle(), None) == INVALID_VALUE
    assert N.lib.b2dgpu_submit(ctx.runtime_handle(), None, None) == INVALID_VALUE
    assert N.lib.b2dgpu_submit(ctx.runtime_handle(), ctx.target_handle(), None) == INVALID_VALUE
    tgt = C.c_void_p()
    assert N.lib.b2dgpu_target_create(ctx.runtime_handle(), 0, 5, 1, C.byref(tgt)) == INVALID_VALUE
    assert N.lib.b2dgpu_target_create(ctx.runtime_handle(), 70000, 5, 1, C.byref(tgt)) == INVALID_VALUE
    assert N.lib.b2dgpu_target_create(ctx.runtime_handle(), 16, 16, 99, C.byref(tgt)) == INVALID_VALUE
    ctx.close()


def test_two_contexts_two_threads_one_runtime(ref, gpu):
    W, H = 400, 300
    rt = gpu.Runtime(device=0)
    results, errors = {}, []

    def work(seed):
        try:
            scene = S.mixed(150, W, H)
            img = gpu.Image(W, H, 1)
            ctx = gpu.Context(img, runtime=rt, command_queue_limit=40)
            scene(gpu, ctx, np.random.default_rng(seed))
            ctx.end()
            results[seed] = img.to_numpy().copy()
            ctx.close()
        except Exception as e:                     # pragma: no cover - reported below
            errors.append(e)
    threads = [threading.Thread(target=work, args=(s,)) for s in (11, 12)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for seed in (11, 12):
        ri, _ = S.draw(ref, S.mixed(150, W, H), W, H, 1, seed)
        n, d = S.channel_diff(ri.to_numpy(), results[seed])
        assert d <= 1, f"seed {seed}: {n} px differ, max {d}"
    rt.close()
