"""GPU tests that need TWO devices (-m gpu; skipped on a single-GPU box): one image split over two ranks with NCCL
(proxtv_b200.distributed.tv1_2d_single_sharded) must reproduce the single-GPU solve."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import proxtv_b200 as ptv
    from oracle import oracle as O
    from proxtv_b200.distributed import tv1_2d_single_sharded
    Y = O.gen_cfg2(512, 768, seed=5, block=16) if rank == 0 else None
    out = tv1_2d_single_sharded(Y, 0.2)
    if rank == 0:
        want = ptv.tv1_2d(Y, 0.2)
        ok = np.abs(out - want).max() <= 1e-9 * np.abs(want).max()
        open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
    dist.barrier(); dist.destroy_process_group()


def test_single_image_over_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 34500 + os.getpid() % 2000, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok")).read() == "1"
