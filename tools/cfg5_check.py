"""Batched float32 DR2_TV (config 5 shape) per image against the float64 reference on the float32-rounded input."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
from oracle import oracle as O
ptv.require_device()
R = O.Ref()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
imgs = np.stack([np.ascontiguousarray(O.gen_cfg2(n, n, seed=s)).astype(np.float32) for s in range(B)])
want = [R.dr2_tv(np.asfortranarray(imgs[s].astype(np.float64)), 0.2, n_threads=16)[0] for s in range(B)]
for eng in ("lane", "chunked"):
    ptv.set_engine(eng)
    got = ptv.tv1_2d_batched(imgs, 0.2)
    single = [ptv.tv1_2d_batched(imgs[s:s + 1], 0.2)[0] for s in range(B)]
    for s in range(B):
        d = np.abs(got[s] - want[s]); sc = np.abs(want[s]).max()
        print(eng, "image", s, "rel err batched %.3e" % (d.max() / sc), "n>2e-5: %d" % int((d > 2e-5 * sc).sum()),
              "single %.3e" % (np.abs(single[s] - want[s]).max() / sc), "batched==single", bool(np.array_equal(got[s], single[s])), flush=True)
