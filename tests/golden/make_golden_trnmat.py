"""Generates tests/golden/trnmat_*.npz.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_trnmat.py

Inputs + outputs of the REFERENCE's own `trans` / `trans_sub` (the producer of `trn_mat_time`,
preprocess_to_trnmat.ipynb cells 7 and 13: the cell sources are exec'd unmodified; the only shim is
`np.int = int`, an alias numpy 2 removed) on small seeded interaction dicts.  They pin
sagnn_b200.data_handler.trans_sub -- interval bucketing, first-occurrence timestamps, timeMat -- bit-exactly.
"""
import io
import contextlib
import json
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
NB = "/root/reference/preprocess_to_trnmat.ipynb"


def reference_functions():
    nb = json.load(open(NB))
    src = {}
    for c in nb["cells"]:
        s = "".join(c["source"])
        if c["cell_type"] == "code" and "def trans_sub(" in s:
            src["trans_sub"] = s
        if c["cell_type"] == "code" and "def trans(" in s:
            src["trans"] = s
    np.int = int                                   # removed alias; the notebook predates numpy 1.24
    ns = {"np": np, "sp": sp, "csr_matrix": sp.csr_matrix}
    exec(src["trans"], ns)                         # defines minn / maxx globals and trans()
    exec(src["trans_sub"], ns)
    return ns


def random_interaction(rng, U, I, n_events, t_lo=1388534400, t_hi=1406073600, repeat=0.3, none_users=0.1):
    """list[U] of None | {item: [timestamps...]}, like the notebook's trnInt."""
    inter = [None if rng.random() < none_users else {} for _ in range(U)]
    users = [u for u in range(U) if inter[u] is not None]
    for _ in range(n_events):
        u = users[rng.integers(len(users))]
        d = inter[u]
        if d and rng.random() < repeat:            # another timestamp on an existing pair
            it = list(d.keys())[rng.integers(len(d))]
        else:
            it = int(rng.integers(I))
        d.setdefault(it, []).append(int(rng.integers(t_lo, t_hi)))
    return inter


def flatten(inter):
    us, its, ts = [], [], []
    for u, d in enumerate(inter):
        if d is None:
            continue
        for it in d:
            for t in d[it]:
                us.append(u); its.append(it); ts.append(t)
    return np.array(us, np.int64), np.array(its, np.int64), np.array(ts, np.int64)


def main():
    ns = reference_functions()
    rng = np.random.default_rng(100)
    cases = {"a_60x40_t5": (60, 40, 700, 5), "b_30x80_t3": (30, 80, 400, 3), "c_25x25_t8": (25, 25, 900, 8),
             "d_10x10_t1": (10, 10, 60, 1)}
    for name, (U, I, n, T) in cases.items():
        inter = random_interaction(rng, U, I, n)
        ns["minn"], ns["maxx"] = 1647180684, 0     # the notebook's initial values (cell 13)
        with contextlib.redirect_stdout(io.StringIO()):
            trn = ns["trans"](inter, U, I)
            sub, tm = ns["trans_sub"](inter, U, I, T)
        u, i, t = flatten(inter)
        out = {"U": U, "I": I, "T": T, "u": u, "i": i, "t": t, "minn": ns["minn"], "maxx": ns["maxx"]}
        trn = sp.csr_matrix(trn); trn.sum_duplicates(); trn.sort_indices()
        out.update(trn_indptr=trn.indptr, trn_indices=trn.indices, trn_data=trn.data)
        tm = sp.csr_matrix(tm); tm.sort_indices()
        out.update(tm_indptr=tm.indptr, tm_indices=tm.indices, tm_data=tm.data)
        for k, m in enumerate(sub):
            m = sp.csr_matrix(m); m.sort_indices()
            out.update({"sub%d_indptr" % k: m.indptr, "sub%d_indices" % k: m.indices, "sub%d_data" % k: m.data})
        np.savez_compressed(os.path.join(HERE, "trnmat_%s.npz" % name), **out)
        print(name, "events", len(t), "nnz", [int(sp.csr_matrix(m).nnz) for m in sub], "timeMat nnz", tm.nnz)


if __name__ == "__main__":
    main()
