"""Seeded synthetic inputs of the BASELINE.json configurations: the generators bench.py, the tools and (through
oracle/oracle.py, which re-exports them) every test share.  Plain numpy; nothing here belongs to the oracle or to the product."""
import numpy as np


def gen_cfg1(n=1_000_000, seed=0):
    rng = np.random.default_rng(seed)
    return np.repeat(rng.normal(0, 2, n // 1000 + 1), 1000)[:n] + rng.normal(0, 0.5, n)


def gen_cfg2(M=4096, N=None, seed=0, block=64):
    N = M if N is None else N
    rng = np.random.default_rng(seed)
    lv = rng.normal(0, 1, (-(-M // block), -(-N // block)))
    img = np.kron(lv, np.ones((block, block)))[:M, :N] + rng.normal(0, 0.3, (M, N))
    return np.asfortranarray(img)


def gen_cfg3(B=65536, L=4096, seed=0):
    rng = np.random.default_rng(seed)
    X = np.repeat(rng.normal(0, 2, (B, -(-L // 64))), 64, axis=1)[:, :L] + rng.normal(0, 0.5, (B, L))
    W = rng.uniform(0.1, 1.0, (B, L - 1))
    return X, W


def gen_cfg4(shape=(512, 512, 256), seed=0, block=16):
    rng = np.random.default_rng(seed)
    lv = rng.normal(0, 1, tuple(-(-s // block) for s in shape))
    V = np.kron(lv, np.ones((block,) * len(shape)))[tuple(slice(0, s) for s in shape)] + rng.normal(0, 0.3, shape)
    return np.asfortranarray(V)
