"""Generates tests/golden/dataset_tiny/ (a dataset directory in the reference's on-disk layout) and
tests/golden/loaddata_tiny.npz by CALLING the reference's own ``DataHandler().LoadData()`` on it
(DataHandler.py:70-129).  Run in the BUILD container only:

    python tests/golden/make_golden_loaddata.py

The directory holds what ``preprocess_to_trnmat.ipynb`` / ``preprocess_to_sequence.ipynb`` write: ``trn_mat_time``
(pickle of ``[trnMat, subMat[T], timeMat]``), ``sequence`` (per-user item lists) and ``tst_int`` (held-out item or
None); the fixture records what the reference keeps after loading: ``args.user`` / ``args.item``, every ``subMat[k]``
(the T interval matrices of the path), the rating matrix it rebuilds from the sequences, ``tstInt`` / ``tstUsrs``,
``maxTime``.  ``tests/test_oracle.py`` checks the product's ``load_trn_mat_time`` mirror against it.
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

from make_golden_model import HERE, REF, load_reference

ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    shim, model, NNs = load_reference()
    import DataHandler as RefDH                      # /root/reference/DataHandler.py
    sys.path.insert(0, ROOT)
    from make_golden_sampler import make_data
    rng = np.random.default_rng(11)
    U, I, T = 37, 29, 4
    seqs, tst, sub, trn_full = make_data(rng, U, I, T, max_len=12)
    trn0 = sp.csr_matrix((trn_full != 0).astype(np.float64))
    time_mat = sp.csr_matrix((U, I), dtype=np.intc)
    for k, m in enumerate(sub):
        if k:
            time_mat = time_mat + ((m != 0).astype(np.intc) * k)
    ddir = os.path.join(HERE, "dataset_tiny")
    os.makedirs(ddir, exist_ok=True)
    with open(os.path.join(ddir, "trn_mat_time"), "wb") as fs:
        pickle.dump([trn0, sub, sp.csr_matrix(time_mat).astype(np.intc)], fs, protocol=4)
    with open(os.path.join(ddir, "sequence"), "wb") as fs:
        pickle.dump(seqs, fs, protocol=4)
    with open(os.path.join(ddir, "tst_int"), "wb") as fs:
        pickle.dump(tst, fs, protocol=4)

    # the reference opens './Datasets/<args.data>/...' relative to the working directory
    work = tempfile.mkdtemp()
    os.makedirs(os.path.join(work, "Datasets"))
    os.symlink(ddir, os.path.join(work, "Datasets", "tiny"))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        RefDH.args.data, RefDH.args.percent = "tiny", 0.0
        h = RefDH.DataHandler()
        h.LoadData()
    finally:
        os.chdir(cwd)
    out = dict(user=RefDH.args.user, item=RefDH.args.item, T=len(h.subMat), maxTime=h.maxTime,
               tstUsrs=np.asarray(h.tstUsrs), tstInt=np.array([-1 if x is None else x for x in h.tstInt], np.int64),
               trn_indptr=h.trnMat.indptr, trn_indices=h.trnMat.indices, trn_data=h.trnMat.data,
               time_indptr=h.timeMat.indptr, time_indices=h.timeMat.indices, time_data=h.timeMat.data,
               seq_ptr=np.concatenate([[0], np.cumsum([len(s) for s in h.sequence])]).astype(np.int64),
               seq_items=np.concatenate(h.sequence).astype(np.int32))
    for k, m in enumerate(h.subMat):
        out["sub%d_indptr" % k], out["sub%d_indices" % k], out["sub%d_data" % k] = m.indptr, m.indices, m.data
        out["sub%d_dtype" % k] = np.array(str(m.dtype))
    np.savez_compressed(os.path.join(HERE, "loaddata_tiny.npz"), **out)
    print("LoadData: user", out["user"], "item", out["item"], "T", out["T"], "nnz", [m.nnz for m in h.subMat],
          "tstUsrs", len(h.tstUsrs), "maxTime", h.maxTime)


if __name__ == "__main__":
    main()
