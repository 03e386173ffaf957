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
