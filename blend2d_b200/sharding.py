"""Multi-GPU sharding of the render path: one process per GPU, no collective while rendering.

Two partitions, both taken from the reference's own threading model:

  band sharding   One big canvas.  Rank r owns a contiguous slab of rows, rounded to the tile height, and replays EVERY
                  command clipped to its slab - exactly what a reference worker does with its bands
                  (raster/workerproc.cpp:260-299, rendercommandprocasync_p.h:153-174).  The only exchange is one gather of
                  the slabs when a single contiguous image is needed (`gather_canvas`), NCCL over NVLink on the GPU box.
  frame sharding  Many independent canvases.  Frame i belongs to rank i mod world; nothing is exchanged.

`torch.distributed` is plumbing here (process group + the final gather); rendering never goes through it.
"""
import ctypes as C

TILE_ROWS = 32         # the largest tile height the compositor picks (csrc/kernels.cu choose_tile_height): slabs are
                       # cut on multiples of it so no tile is shared by two ranks whatever height a launch chooses


def slab_table(height, world, align=TILE_ROWS):
    """[(y0, y1)] per rank: contiguous, disjoint, covering [0, height), every boundary a multiple of `align`.
    Ranks that would get nothing (more ranks than row groups) receive an empty slab (y0 == y1): such a rank creates no
    target (b2dgpu_target_create_slab rejects y0 >= y1) and contributes nothing to the gather - see `StripeGather`."""
    if height <= 0 or world <= 0:
        raise ValueError("height and world must be positive")
    groups = (height + align - 1) // align
    table, g0 = [], 0
    for r in range(world):
        g1 = g0 + groups // world + (1 if r < groups % world else 0)
        table.append((min(g0 * align, height), min(g1 * align, height)))
        g0 = g1
    return table


def slab_rows(height, world, rank, align=TILE_ROWS):
    return slab_table(height, world, align)[rank]


def stripe_table(height, world, stripes_per_rank, align=TILE_ROWS):
    """Interleaved band ownership (the reference hands bands to workers round robin for the same reason,
    raster/workerproc.cpp:260-299): the canvas is cut into world * stripes_per_rank tile-aligned stripes and stripe j
    belongs to rank j mod world, so every rank gets rows from all over the canvas and uneven coverage averages out.
    Returns [(y0, y1)] for all stripes in canvas order."""
    return slab_table(height, world * stripes_per_rank, align)


def stripes_of(rank, world, stripes_per_rank, height, align=TILE_ROWS):
    t = stripe_table(height, world, stripes_per_rank, align)
    return [t[j] for j in range(rank, len(t), world)]


def gather_stripes(local_stripes, height, stripes_per_rank, dst=0, group=None):
    """Gathers interleaved stripes into the full image on rank `dst`.
    local_stripes: list of tensors [rows of stripe, row_bytes], this rank's stripes in canvas order."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    table = stripe_table(height, world, stripes_per_rank)
    tallest = max(b - a for a, b in table)
    row_bytes = local_stripes[0].shape[1]
    send = torch.zeros((stripes_per_rank, tallest, row_bytes), dtype=local_stripes[0].dtype, device=local_stripes[0].device)
    mine = [table[j] for j in range(rank, len(table), world)]
    for i, ((a, b), t) in enumerate(zip(mine, local_stripes)):
        send[i, : b - a].copy_(t[: b - a])
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for j, (a, b) in enumerate(table):
        parts.append(recv[j % world][j // world, : b - a])
    return torch.cat(parts, dim=0)


class StripeGather:
    """The one exchange of a band-sharded frame: every stripe goes STRAIGHT into its rows of the final image on rank
    `dst` - no padded staging buffer, no concatenation - and stripe by stripe, so the transfer of stripe i overlaps the
    render of stripe i + 1:

        g = StripeGather(height, stripes_per_rank, rank, world, device)
        g.begin()
        for i, stripe in enumerate(my_stripes):      # canvas order
            render(stripe)                           # enqueued on `stream`
            g.stripe_ready(i, rows_of(stripe), stream)
        image = g.finish(stream)                     # [height, row_bytes] on dst, None elsewhere

    Point-to-point NCCL send/recv (gloo on CPU): the owner of stripe j = rank j mod world sends it, `dst` receives into
    image[y0:y1].  `dst` posts the receives of round i (one stripe from every peer) as one group when its own stripe i
    is ready, so the peers' transfers run concurrently over NVSwitch.  Empty stripes (more stripes than tile rows) are
    skipped on both sides."""

    def __init__(self, height, stripes_per_rank, rank, world, device=None, dst=0, group=None):
        self.height, self.k, self.rank, self.world, self.dst, self.group, self.device = height, stripes_per_rank, rank, world, dst, group, device
        self.table = stripe_table(height, world, stripes_per_rank)
        self.full = None
        self.works = []
        self._copy_stream = None

    def _rows(self, r, i):
        return self.table[r + i * self.world]

    def begin(self):
        self.works = []

    def stripe_ready(self, i, tensor, stream=None):
        import contextlib
        import torch
        import torch.distributed as dist
        y0, y1 = self._rows(self.rank, i)
        cuda = tensor.is_cuda
        ctx = torch.cuda.stream(stream) if (cuda and stream is not None) else contextlib.nullcontext()
        with ctx:
            if self.rank != self.dst:
                if y1 > y0:
                    t = tensor[: y1 - y0]
                    if not t.is_contiguous():
                        t = t.contiguous()                  # padded row stride: one staging copy
                    self.works.append((dist.isend(t, self.dst, group=self.group), t))
                return
            if self.full is None:
                self.full = torch.empty((self.height, tensor.shape[1]), dtype=tensor.dtype, device=tensor.device)
            ops = []
            for r in range(self.world):
                if r == self.dst:
                    continue
                a, b = self._rows(r, i)
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, self.full[a:b], r, group=self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    self.works.append((w, None))
            if y1 > y0:
                if cuda:
                    if self._copy_stream is None:
                        self._copy_stream = torch.cuda.Stream(device=tensor.device)
                    self._copy_stream.wait_stream(stream if stream is not None else torch.cuda.current_stream())
                    with torch.cuda.stream(self._copy_stream):
                        self.full[y0:y1].copy_(tensor[: y1 - y0], non_blocking=True)
                else:
                    self.full[y0:y1].copy_(tensor[: y1 - y0])

    def finish(self, stream=None):
        import contextlib
        import torch
        cuda = self.device is not None and torch.device(self.device).type == "cuda"
        ctx = torch.cuda.stream(stream) if (cuda and stream is not None) else contextlib.nullcontext()
        with ctx:
            for w, _keep in self.works:
                w.wait()
            if cuda and self._copy_stream is not None:
                (stream if stream is not None else torch.cuda.current_stream()).wait_stream(self._copy_stream)
        self.works = []
        return self.full if self.rank == self.dst else None

    def run(self, local_stripes, stream=None):
        self.begin()
        for i, t in enumerate(local_stripes):
            self.stripe_ready(i, t, stream)
        return self.finish(stream)


def frames_of(rank, world, frame_count):
    """Indices of the frames rank `rank` renders (round robin, like scene i -> GPU i mod G)."""
    return range(rank, frame_count, world)


def canvas_tensor(ctx):
    """The context's device canvas slab as a torch uint8 tensor [padded rows, stride bytes] (no copy).
    Rows beyond the slab height and bytes beyond width * bpp are padding."""
    import torch
    from . import _native as N
    ptr, stride, pw, ph = C.c_void_p(), C.c_ssize_t(), C.c_int32(), C.c_int32()
    N.check(N.lib.b2dgpu_target_device_view(ctx.target_handle(), C.byref(ptr), C.byref(stride), C.byref(pw), C.byref(ph)),
            "b2dgpu_target_device_view")

    class _Mem:
        pass
    m = _Mem()
    m.__cuda_array_interface__ = {"shape": (ph.value, stride.value), "typestr": "|u1", "data": (ptr.value, False), "version": 2}
    return torch.as_tensor(m, device=torch.device("cuda", torch.cuda.current_device()))


def gather_canvas(local_rows, height, dst=0, group=None):
    """Gathers per-rank row slabs into the full image on rank `dst`.

    local_rows: torch tensor [rows of this rank's slab, row_bytes] (uint8; CUDA with NCCL, CPU with gloo).
    Returns the [height, row_bytes] tensor on `dst`, None elsewhere.  Slabs are padded to the tallest slab for the
    collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    table = slab_table(height, world)
    y0, y1 = table[rank]
    if local_rows.shape[0] < y1 - y0:
        raise ValueError("local slab has fewer rows than the partition assigns to this rank")
    tallest = max(b - a for a, b in table)
    send = torch.zeros((tallest, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    send[: y1 - y0].copy_(local_rows[: y1 - y0])
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([recv[r][: b - a] for r, (a, b) in enumerate(table)], dim=0)
