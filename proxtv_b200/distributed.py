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


def _peer(group, r):
    """Global rank of group member r (P2POp peers and broadcast sources are global ranks)."""
    import torch.distributed as dist
    return r if group is None else dist.get_global_rank(group, r)


def tv1_2d_batched_sharded(x, w, max_iters=0, src=0, group=None, solver=None, device=None, pieces=1, timings=None):
    """DR2_TV on every image of x (B, H, W), sharded over the ranks of `group`.

    x       on rank `src` (a rank index inside `group`): torch tensor or numpy array (float32 / float64); ignored elsewhere
    solver  callable(local (b, H, W) tensor, w, max_iters) -> tensor; default: the CUDA path of this package
    pieces  > 1: every rank's slab is cut into that many pieces and the transfers are pipelined with the solves -- piece k + 1
            is received (and the result of piece k - 1 sent back) while piece k is being solved, so that only the first
            piece's scatter and the last piece's gather are exposed
    timings optional dict: receives the wall time of the phases on this rank ('scatter', 'solve', 'gather' seconds; with
            pieces > 1 the phases overlap and only 'total' is meaningful)
    Returns the (B, H, W) result on rank `src`, None on the other ranks.
    """
    import time
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group); world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    if pieces > 1:
        return _sharded_pipelined(x, w, max_iters, src, group, solver, device, pieces, timings)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = [None]
    if rank == src:
        xt = torch.as_tensor(x)
        assert xt.dim() == 3 and xt.dtype in (torch.float32, torch.float64)
        meta = [(tuple(xt.shape), str(xt.dtype).split(".")[-1])]
    dist.broadcast_object_list(meta, src=_peer(group, src), group=group)
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
                ops.append(dist.P2POp(dist.isend, xs[a:b].contiguous(), _peer(group, r), group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.irecv, local, _peer(group, src), group))
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
                ops.append(dist.P2POp(dist.irecv, out[a:b], _peer(group, r), group))
    elif hi > lo:
        ops.append(dist.P2POp(dist.isend, res, _peer(group, src), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def _sharded_pipelined(x, w, max_iters, src, group, solver, device, pieces, timings):
    """Software-pipelined variant: transfers of neighbouring pieces overlap the solve of the current one."""
    import time
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
    dist.broadcast_object_list(meta, src=_peer(group, src), group=group)
    (B, H, W), dname = meta[0]
    dtype = getattr(torch, dname)
    bounds = slab_bounds(B, world)
    if solver is None:
        from . import tv1_2d_batched
        solver = lambda t, ww, it: tv1_2d_batched(t, ww, max_iters=it)      # noqa: E731
    # piece p of rank r: images [a + cut(p), a + cut(p + 1)) of its slab [a, b)
    def cut(a, b, p):
        return a + (b - a) * p // pieces
    sync = (lambda: torch.cuda.synchronize()) if device.type == "cuda" else (lambda: None)
    t0 = time.perf_counter()
    xs = xt.to(device).contiguous() if rank == src else None
    out = torch.empty((B, H, W), dtype=dtype, device=device) if rank == src else None
    lo, hi = bounds[rank]
    bufs = [None] * pieces; res = [None] * pieces
    pending = []

    def post_scatter(p):
        ops = []
        if rank == src:
            for r, (a, b) in enumerate(bounds):
                pa, pb = cut(a, b, p), cut(a, b, p + 1)
                if r == src:
                    bufs[p] = xs[pa:pb]
                elif pb > pa:
                    ops.append(dist.P2POp(dist.isend, xs[pa:pb], _peer(group, r), group))
        else:
            pa, pb = cut(lo, hi, p), cut(lo, hi, p + 1)
            bufs[p] = torch.empty((pb - pa, H, W), dtype=dtype, device=device)
            if pb > pa:
                ops.append(dist.P2POp(dist.irecv, bufs[p], _peer(group, src), group))
        return dist.batch_isend_irecv(ops) if ops else []

    def post_gather(p):
        ops = []
        if rank == src:
            for r, (a, b) in enumerate(bounds):
                pa, pb = cut(a, b, p), cut(a, b, p + 1)
                if r == src:
                    out[pa:pb].copy_(res[p])
                elif pb > pa:
                    ops.append(dist.P2POp(dist.irecv, out[pa:pb], _peer(group, r), group))
        elif res[p].shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, res[p], _peer(group, src), group))
        return dist.batch_isend_irecv(ops) if ops else []

    reqs = {0: post_scatter(0)}
    for p in range(pieces):
        for q in reqs.pop(p):
            q.wait()
        if p + 1 < pieces:
            reqs[p + 1] = post_scatter(p + 1)                # next piece travels while this one is solved
        r_ = solver(bufs[p], w, max_iters) if bufs[p].shape[0] > 0 else bufs[p]
        res[p] = torch.as_tensor(r_).to(device=device, dtype=dtype).contiguous()
        sync()                                               # the result must be complete before it is sent
        pending.append(post_gather(p))
        bufs[p] = None
    for g in pending:
        for q in g:
            q.wait()
    sync()
    if timings is not None:
        timings["total"] = time.perf_counter() - t0
    return out
