"""Generates tests/golden/*.npz.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

index_*.npz   -- inputs + outputs of the REFERENCE's own DataHandler.transToLsts /
                 DataHandler.transpose (imported from /root/reference, unmodified) on small
                 seeded matrices: these pin the oracle's and the device plan's index order,
                 fallback edge and int32-truncated "normalised" data bit-exactly.
prop_*.npz    -- propagation inputs + fp64 outputs produced by oracle/propagate_oracle.py
                 (NOT by the reference: TF 1.14 is not importable; parity for the float path
                 is unpinned, these only freeze the oracle against regressions).
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def ref_handler():
    sys.argv = ["x"]            # Params.py parses argv at import
    sys.path.insert(0, REF)
    import DataHandler          # the reference module, unmodified
    return DataHandler


def small_matrices():
    out = {}
    rng = np.random.default_rng(100)

    def rand(U, I, nnz, dtype=np.intc, ts=True):
        keys = rng.choice(U * I, size=nnz, replace=False)
        r, c = keys // I, keys % I
        v = rng.integers(1388534400, 1406073600, size=nnz) if ts else rng.integers(1, 5, size=nnz)
        return sp.csr_matrix((v.astype(dtype), (r, c)), shape=(U, I))

    out["ts_40x30"] = rand(40, 30, 150)
    out["ts_7x90"] = rand(7, 90, 60)
    out["small_vals_25x25"] = rand(25, 25, 80, ts=False)
    out["float_counts_30x20"] = rand(30, 20, 100, dtype=np.float64, ts=False)
    out["single_edge_5x6"] = sp.csr_matrix((np.array([1400000000], np.intc), ([3], [2])), shape=(5, 6))
    out["empty_6x4"] = sp.csr_matrix((6, 4), dtype=np.intc)
    m = rand(300, 200, 2500)     # gaps: last populated row far from the end
    m = sp.csr_matrix(m.multiply(sp.csr_matrix(np.arange(300)[:, None] < 150)).astype(np.intc))
    m.eliminate_zeros()
    out["gap_300x200"] = m
    return out


def main():
    D = ref_handler()
    for name, m in small_matrices().items():
        rec = dict(shape=np.array(m.shape), indptr=m.indptr, indices=m.indices, data=m.data)
        for norm in (False, True):
            idx, dat, shp = D.transToLsts(m, norm=norm)
            tidx, tdat, tshp = D.transToLsts(D.transpose(m), norm=norm)
            tag = "norm" if norm else "raw"
            rec.update({f"adj_idx_{tag}": idx, f"adj_data_{tag}": dat, f"adj_shape_{tag}": np.array(shp),
                        f"tp_idx_{tag}": tidx, f"tp_data_{tag}": tdat, f"tp_shape_{tag}": np.array(tshp)})
        t = D.transpose(m)
        rec.update(tp_indptr=t.indptr, tp_indices=t.indices, tp_data=t.data,
                   rowsum=np.asarray(np.sum(m, axis=1)).reshape(-1), colsum=np.asarray(np.sum(m, axis=0)).reshape(-1))
        np.savez_compressed(os.path.join(HERE, f"index_{name}.npz"), **rec)
        print("index", name, m.shape, m.nnz)

    sys.path.insert(0, ROOT)
    from oracle import propagate_oracle as po
    rng = np.random.default_rng(7)
    for name, (T, U, I, d, L, dens, leaky) in {
        "t3_l2_d64": (3, 60, 45, 64, 2, 0.08, 0.5),
        "t2_l3_d32": (2, 50, 70, 32, 3, 0.05, 0.5),
        "t1_l1_d128_leaky01": (1, 33, 21, 128, 1, 0.2, 0.1),
    }.items():
        adj, tp, mats = [], [], []
        for k in range(T):
            m = sp.random(U, I, density=dens, random_state=1000 + k, format="csr")
            m.data[:] = 1
            m = m.astype(np.intc)
            mats.append(m)
            adj.append(D.transToLsts(m)[0])
            tp.append(D.transToLsts(D.transpose(m))[0])
        uE = rng.normal(size=(T, U, d)).astype(np.float32)
        iE = rng.normal(size=(T, I, d)).astype(np.float32)
        gU = rng.normal(size=(T, U, d)).astype(np.float32)
        gI = rng.normal(size=(T, I, d)).astype(np.float32)
        uv, iv, du, di = po.propagate(adj, tp, uE, iE, gU, gI, L, leaky, np.float64)
        rec = dict(T=T, U=U, I=I, d=d, L=L, leaky=leaky, uE=uE, iE=iE, gU=gU, gI=gI,
                   user_vec=uv, item_vec=iv, dU=du, dI=di)
        for k in range(T):
            rec[f"adj{k}"] = adj[k]
            rec[f"tp{k}"] = tp[k]
        np.savez_compressed(os.path.join(HERE, f"prop_{name}.npz"), **rec)
        print("prop", name)


if __name__ == "__main__":
    main()
