"""ONE image over several GPUs (tv1_2d_single_sharded) under torchrun: result against the single-GPU solve, time per solve.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/split_check.py [size]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
from proxtv_b200.distributed import tv1_2d_single_sharded
import synth_inputs as O
local = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for size in ([int(a) for a in sys.argv[1:]] or [1024, 4096]):
    x = want = None
    if rank == 0:
        Y = O.gen_cfg2(size, size, seed=0)
        x = torch.tensor(np.ascontiguousarray(Y.T), device="cuda").t()          # column-major view
        want = ptv.tv1_2d(x, 0.2); torch.cuda.synchronize()
        t0 = time.perf_counter(); ptv.tv1_2d(x, 0.2); torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    for f32 in (False, True):
        xin = x.float() if (rank == 0 and f32) else x
        tm = {}
        out = tv1_2d_single_sharded(xin, 0.2, timings=tm)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = tv1_2d_single_sharded(xin, 0.2, timings=tm); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if rank == 0:
            err = ((out.double() - want).abs().max() / want.abs().max()).item()
            print("%d x %d %s over %d GPUs: rel err vs single GPU %.2e | %.2f ms per solve (%.2f ms in exchanges on rank 0) vs %.2f ms on one GPU"
                  % (size, size, "f32" if f32 else "f64", world, err, dt * 1e3, tm["exchange"] * 1e3, t1 * 1e3), flush=True)
            assert err <= (5e-5 if f32 else 1e-9)
dist.barrier(); dist.destroy_process_group()
