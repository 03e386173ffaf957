"""torchrun --nproc-per-node N tools/shard_check.py : config-5 style batch sharded over N GPUs (NCCL scatter/gather), checked
against the single-GPU result on rank 0; prints aggregate Mpix/s."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
from proxtv_b200.distributed import tv1_2d_batched_sharded
import synth_inputs as O
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
x = None
if rank == 0:
    x = torch.stack([torch.from_numpy(np.ascontiguousarray(O.gen_cfg2(H, H, seed=s).astype(np.float32))) for s in range(B)]).cuda()
out = tv1_2d_batched_sharded(x, 0.2)            # warm-up + correctness
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
out = tv1_2d_batched_sharded(x, 0.2)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    ref = ptv.tv1_2d_batched(x, 0.2)
    print("sharded over %d GPUs: %d x %dx%d f32 in %.1f ms (incl. scatter/gather) = %.1f Mpix/s; equal to single-GPU result: %s"
          % (world, B, H, H, dt * 1e3, B * H * H / dt / 1e6, bool(torch.equal(out, ref))))
dist.destroy_process_group()
