"""CPU test (-m "not gpu"), world_size 2 over gloo: the batch-sharding driver (proxtv_b200/distributed.py) scatters,
solves and gathers correctly, including uneven splits and an idle rank.  The per-rank solver is injected (the oracle port,
image by image) because no GPU exists here; on the GPU box the default solver is the CUDA path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, tmp, pieces=1):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from proxtv_b200.distributed import slab_bounds, tv1_2d_batched_sharded
    P = O.Port()

    def solver(t, w, it):
        return torch.stack([torch.from_numpy(np.ascontiguousarray(P.dr2_tv(img.numpy(), w, maxit=it)[0])) for img in t])

    imgs = None
    if rank == 0:
        imgs = np.stack([np.ascontiguousarray(O.gen_cfg2(24, 20, seed=s, block=4)) for s in range(batch)])
    out = tv1_2d_batched_sharded(imgs, 0.2, max_iters=5, src=0, solver=solver, pieces=pieces)
    if rank == 0:
        want = np.stack([P.dr2_tv(im, 0.2, maxit=5)[0] for im in imgs])
        ok = out.shape == want.shape and np.array_equal(out.numpy(), want)
        b = slab_bounds(batch, world)
        ok = ok and b[0][0] == 0 and b[-1][1] == batch and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        open(os.path.join(tmp, "ok_%d" % batch), "w").write("1" if ok else "0")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 1])       # uneven split; fewer images than ranks (rank 1 idles)
def test_batch_sharding_over_gloo(batch, tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + batch) % 2000
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok_%d" % batch)).read() == "1"


def test_batch_sharding_pipelined_over_gloo(tmp_path):
    """pieces > 1: the transfers of neighbouring pieces overlap the solves; same result, uneven pieces included."""
    import torch.multiprocessing as mp
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, 7, str(tmp_path), 3), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok_7")).read() == "1"


def test_slab_bounds():
    sys.path.insert(0, ROOT)
    from proxtv_b200.distributed import slab_bounds
    assert slab_bounds(1024, 8) == [(128 * r, 128 * (r + 1)) for r in range(8)]
    assert slab_bounds(3, 2) == [(0, 2), (2, 3)] and slab_bounds(0, 2) == [(0, 0), (0, 0)]


def _split_worker(rank, world, port, tmp):
    """tv1_2d_single_sharded over gloo with oracle-backed passes (no GPU here): ONE image, column / row slabs, all-to-all twice per
    iteration; the result must equal the oracle's DR2_TV (only the summation order of the start value differs)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from proxtv_b200.distributed import tv1_2d_single_sharded
    P = O.Port()

    def prox_cols(t_c, w):                          # [n][M]: contiguous fibers
        return torch.from_numpy(np.stack([P.tv1_linearized(f.numpy(), w) for f in t_c]))

    def row_pass(final, y_r, x1_r, t_r, w):         # [N][m]: fiber of row i = column i of the tensor
        d = t_r - x1_r
        u = y_r - d if final else y_r - (2.0 * d - t_r)
        x2 = torch.from_numpy(np.stack([P.tv1_linearized(np.ascontiguousarray(u[:, i].numpy()), w) for i in range(u.shape[1])], axis=1))
        return x2 if final else d + x2

    ok = True
    for (M, N, it) in ((24, 20, 5), (16, 36, 0)):
        Y = O.gen_cfg2(M, N, seed=3 + M, block=4) if rank == 0 else None
        tm = {}
        out = tv1_2d_single_sharded(Y, 0.2, max_iters=it, src=0, passes=(prox_cols, row_pass), timings=tm)
        if rank == 0:
            want = P.dr2_tv(Y, 0.2, maxit=it)[0] if it else P.dr2_tv(Y, 0.2)[0]
            ok = ok and out.shape == want.shape and np.abs(out - want).max() <= 1e-12 * np.abs(want).max() and "exchange" in tm
        else:
            ok = ok and out is None
    if rank == 0:
        open(os.path.join(tmp, "split_ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_single_image_split_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 33500 + os.getpid() % 2000
    mp.spawn(_split_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "split_ok")).read() == "1"
