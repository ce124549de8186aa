"""Host logic of the multi-GPU path on CPU: partition tables, and a world_size-2 gloo run of the band-sharded
gather and of frame sharding.  The ranks render with the host simulator; on the GPU box the same functions run over NCCL
(bench.py --gpus N, tests under -m gpu for the slab targets)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blend2d_b200 import sharding as SH
from tests import scenes as S


@pytest.mark.parametrize("height,world", [(2160, 1), (2160, 2), (2160, 8), (16384, 8), (600, 7), (13, 4), (8, 8), (1, 2)])
def test_slab_table_partitions_rows(height, world):
    t = SH.slab_table(height, world)
    assert len(t) == world and t[0][0] == 0 and t[-1][1] == height
    for (a0, a1), (b0, b1) in zip(t, t[1:]):
        assert a1 == b0 and a0 <= a1
    assert all(a % SH.TILE_ROWS == 0 for a, _ in t if a < height)
    sizes = [b - a for a, b in t]
    assert max(sizes) - min(sizes) < 2 * SH.TILE_ROWS or height < world * SH.TILE_ROWS   # one group, plus the clipped last slab


def test_frames_round_robin():
    owned = [list(SH.frames_of(r, 4, 10)) for r in range(4)]
    assert sorted(sum(owned, [])) == list(range(10))
    assert owned[1] == [1, 5, 9]


def test_stripes_interleave_and_cover():
    t = SH.stripe_table(2160, 4, 3)
    assert len(t) == 12 and t[0][0] == 0 and t[-1][1] == 2160
    owned = [SH.stripes_of(r, 4, 3, 2160) for r in range(4)]
    assert sorted(sum(owned, [])) == t
    assert owned[1] == [t[1], t[5], t[9]]


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


W, H = 200, 90


def full_render(seed):
    from tests import hostsim
    return hostsim.draw(S.mixed(40, W, H), W, H, 1, seed)


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # band sharding: every rank replays the whole scene and keeps its own rows (the GPU clips to the slab in
        # k_finalize_commands; the simulator renders the full canvas, so the slab is cut out here)
        y0, y1 = SH.slab_rows(H, world, rank)
        slab = torch.from_numpy(full_render(7)[y0:y1].copy().view(np.uint8))
        canvas = SH.gather_canvas(slab, H, dst=0)
        # interleaved stripes: two per rank
        full_img = full_render(7)
        mine = [torch.from_numpy(full_img[a:b].copy().view(np.uint8)) for a, b in SH.stripes_of(rank, world, 2, H)]
        striped = SH.gather_stripes(mine, H, 2, dst=0)
        if rank == 0:
            assert np.array_equal(striped.numpy().view(np.uint32), full_img)
        # the same exchange stripe by stripe, straight into the final image rows (what bench.py's band-sharded leg uses);
        # 3 stripes per rank on 90 rows = more stripes than tile rows: the empty ones are skipped on both sides
        for k in (2, 3):
            mine_k = [torch.from_numpy(full_img[a:b].copy().view(np.uint8)) for a, b in SH.stripes_of(rank, world, k, H)]
            g = SH.StripeGather(H, k, rank, world)
            g.begin()
            for i, t in enumerate(mine_k):
                g.stripe_ready(i, t)
            direct = g.finish()
            if rank == 0:
                assert np.array_equal(direct.numpy().view(np.uint32), full_img)
            else:
                assert direct is None
        # frame sharding: independent frames, only a checksum of checksums is reduced for the report
        sums = torch.zeros(6, dtype=torch.int64)
        for i in SH.frames_of(rank, world, 6):
            sums[i] = int(full_render(100 + i).astype(np.uint64).sum() % (2 ** 31))
        dist.all_reduce(sums)
        if rank == 0:
            np.save(os.path.join(out_dir, "canvas.npy"), canvas.numpy())
            np.save(os.path.join(out_dir, "sums.npy"), sums.numpy())
        else:
            assert canvas is None
    finally:
        dist.destroy_process_group()


def test_world2_gloo_band_gather_and_frame_sharding(tmp_path):
    world = 2
    from tests import hostsim
    hostsim.lib()                 # build the simulator once here: the two workers must not run `make` on it concurrently
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    canvas = np.load(tmp_path / "canvas.npy").view(np.uint32)
    assert np.array_equal(canvas, full_render(7))
    sums = np.load(tmp_path / "sums.npy")
    want = [int(full_render(100 + i).astype(np.uint64).sum() % (2 ** 31)) for i in range(6)]
    assert sums.tolist() == want


@pytest.mark.gpu
def test_slab_contexts_cover_the_canvas(ref, gpu):
    """Two slab contexts on one device replay the same scene; each writes only its own rows of the host image."""
    w, h = 300, 203
    scene = S.mixed(120, w, h)
    ri, _ = S.draw(ref, scene, w, h, 1, 5)
    img = gpu.Image(w, h, 1)
    for rank in range(3):
        ctx = gpu.Context(img, slab=SH.slab_rows(h, 3, rank))
        scene(gpu, ctx, np.random.default_rng(5))
        ctx.end(); ctx.close()
    n, d = S.channel_diff(ri.to_numpy(), img.to_numpy())
    assert d <= 1, f"{n} pixels differ, max {d}"
