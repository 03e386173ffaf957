"""One process, two devices in turn: results on cuda:0 and cuda:1 must be identical (per-device workspaces / streams / graphs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
from oracle import oracle as O
assert torch.cuda.device_count() >= 2
Y = np.ascontiguousarray(O.gen_cfg2(1536, 1024, seed=3)); Ys = np.ascontiguousarray(O.gen_cfg2(200, 136, seed=4, block=8))
want = O.Port().dr2_tv(Ys, 0.3)[0]
res = {}
for rnd in range(2):
    for d in (0, 1, 1, 0):
        dev = torch.device("cuda:%d" % d)
        a = ptv.tv1_2d(torch.tensor(Y.T.copy(), device=dev).t(), 0.2)          # column-major view: pipelined + graph path
        b = ptv.tv1_2d(torch.tensor(Ys, device=dev), 0.3)                      # small row-major: serial path
        torch.cuda.synchronize(dev)
        res.setdefault("big", []).append(a.cpu().numpy()); res.setdefault("small", []).append(b.cpu().numpy())
ok_big = all(np.array_equal(res["big"][0], r) for r in res["big"]); ok_small = all(np.array_equal(res["small"][0], r) for r in res["small"])
err = np.abs(res["small"][0] - want).max() / np.abs(want).max()
print("devices agree (big, small):", ok_big, ok_small, " small vs oracle rel err %.2e" % err)
assert ok_big and ok_small and err < 1e-9
