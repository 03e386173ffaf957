"""Multi-GPU driver for the one case the path shards: a BATCH of independent images (BASELINE config 5).

One process per GPU (torchrun); rank `src` holds the batch, every rank solves a contiguous slab of images with the
single-GPU path (`proxtv_b200.tv1_2d_batched`), and the slabs are gathered back on `src`.  There is no communication
inside a solve: a single image cannot be split without an all-to-all transpose between the two passes of every
iteration, which the north star excludes.  Collectives: one scatter and one gather of image slabs
(point-to-point sends grouped with `batch_isend_irecv`, so uneven slabs need no padding) over NCCL/NVLink, or gloo on CPU
for the host-logic tests.
"""
import numpy as np


def slab_bounds(batch, world):
    """Contiguous, balanced split of `batch` images over `world` ranks: returns [(start, stop)] per rank."""
    base, extra = divmod(int(batch), int(world))
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def tv1_2d_batched_sharded(x, w, max_iters=0, src=0, group=None, solver=None, device=None):
    """DR2_TV on every image of x (B, H, W), sharded over the ranks of `group`.

    x       on rank `src`: torch tensor or numpy array (float32 / float64); ignored elsewhere (pass None)
    solver  callable(local (b, H, W) tensor, w, max_iters) -> tensor; default: the CUDA path of this package
    Returns the (B, H, W) result on rank `src`, None on the other ranks.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group); world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = [None]
    if rank == src:
        xt = torch.as_tensor(x)
        assert xt.dim() == 3 and xt.dtype in (torch.float32, torch.float64)
        meta = [(tuple(xt.shape), str(xt.dtype).split(".")[-1])]
    dist.broadcast_object_list(meta, src=src, group=group)
    (B, H, W), dname = meta[0]
    dtype = getattr(torch, dname)
    bounds = slab_bounds(B, world)
    lo, hi = bounds[rank]
    local = torch.empty((hi - lo, H, W), dtype=dtype, device=device)

    # ---- scatter ----
    ops = []
    if rank == src:
        xs = xt.to(device).contiguous()
        for r, (a, b) in enumerate(bounds):
            if r == src:
                local.copy_(xs[a:b])
            elif b > a:
                ops.append(dist.P2POp(dist.isend, xs[a:b].contiguous(), r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, local, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    # ---- local solve (no communication) ----
    if solver is None:
        from . import tv1_2d_batched
        solver = lambda t, ww, it: tv1_2d_batched(t, ww, max_iters=it)      # noqa: E731
    res = solver(local, w, max_iters) if hi > lo else local
    res = torch.as_tensor(res).to(device=device, dtype=dtype).contiguous()

    # ---- gather ----
    ops = []; out = None
    if rank == src:
        out = torch.empty((B, H, W), dtype=dtype, device=device)
        for r, (a, b) in enumerate(bounds):
            if r == src:
                out[a:b].copy_(res)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.isend, res, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
