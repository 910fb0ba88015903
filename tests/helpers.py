"""Shared test helpers (graph builders, oracle adapters)."""
import numpy as np
import scipy.sparse as sp

from oracle import propagate_oracle as po


def random_interval_mats(T, U, I, nnz, seed=0, timestamps=True):
    rng = np.random.default_rng(seed)
    mats = []
    for _ in range(T):
        n = min(nnz, U * I)
        keys = rng.choice(U * I, size=n, replace=False)
        r, c = keys // I, keys % I
        v = rng.integers(1388534400, 1406073600, size=n) if timestamps else np.ones(n)
        mats.append(sp.csr_matrix((v.astype(np.intc), (r, c)), shape=(U, I)))
    return mats


def adj_lists(mats):
    adj = [po.trans_to_lsts(m)[0] for m in mats]
    tp = [po.trans_to_lsts(po.transpose(m))[0] for m in mats]
    return adj, tp


def random_tables(T, U, I, d, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    return tuple((rng.standard_normal((T, n, d)) * scale).astype(np.float32) for n in (U, I, U, I))


def decode_gpu_masks(masks, T, U, I, d, L):
    """uint8 device buffer written by sagnn_propagate_fwd -> uint8 [T, L, (U+I)*d] in the C
    oracle's layout (1 = gradient passes unscaled).  Device layout per layer: bytes
    [T,U,d/4] then [T,I,d/4]; bit i (i < 4) of byte j of a row is element 4*j + i."""
    raw = masks.detach().cpu().numpy().view(np.uint8)
    bpr = d // 4
    per_layer = T * (U + I) * bpr
    out = np.empty((T, L, (U + I) * d), dtype=np.uint8)
    for l in range(L):
        lay = raw[l * per_layer:(l + 1) * per_layer]
        bits = np.unpackbits(lay.reshape(-1, 1), axis=1, bitorder="little")[:, :4].reshape(-1)
        out[:, l, :U * d] = bits[:T * U * d].reshape(T, U * d)
        out[:, l, U * d:] = bits[T * U * d:].reshape(T, I * d)
    return out
