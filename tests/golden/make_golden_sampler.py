"""Generates tests/golden/sampler_*.npz by CALLING the reference's own samplers.  Run in the BUILD container only:

    python tests/golden/make_golden_sampler.py

``import model`` imports /root/reference/model.py unmodified (over the TF stand-in, which the samplers never touch)
and the fixtures record, for seeded numpy / ``random`` generators (main.py:21-22 seeds both with 100):

  * ``Recommender.sampleSslBatch(self, batIds, labelMat)``                       model.py:304-339
  * ``Recommender.sampleTrainBatch(self, batIds, labelMat, timeMat, n)``         model.py:252-302, which calls the
    reference's ``negSamp`` (DataHandler.py:28-41)
  * an epoch-like sequence -- ``np.random.permutation`` (model.py:342), then for every batch sampleTrainBatch followed
    by sampleSslBatch (model.py:352-356) -- so the fixtures also pin how far each call advances the two streams

with inputs (interval matrices, label matrix, sequences, tstInt), outputs and the generator states afterwards.
"""
import os
import random

import numpy as np
import scipy.sparse as sp

from make_golden_model import HERE, load_reference


class _Obj:
    pass


def make_data(rng, U, I, T, min_len=3, max_len=40, stored_zero=False):
    """Per-user time-ordered item sequences -> interval matrices (values = timestamps), the train matrix (sequence
    without its last item... the reference's trnMat holds the training interactions), tstInt with None entries."""
    seqs, tst = [], []
    rows, cols, vals, ks = [], [], [], []
    for u in range(U):
        n = int(rng.integers(min_len, min(max_len, I // 2) + 1))
        items = rng.choice(I, size=n, replace=False)
        times = np.sort(rng.integers(1, 10 ** 6, size=n))
        seqs.append([int(x) for x in items])
        tst.append(None if rng.random() < 0.3 else int(rng.integers(0, I)))
        for j in range(n):
            rows.append(u); cols.append(int(items[j])); vals.append(int(times[j])); ks.append(min(T - 1, j * T // n))
    rows, cols, vals, ks = map(np.asarray, (rows, cols, vals, ks))
    sub = [sp.csr_matrix((vals[ks == k], (rows[ks == k], cols[ks == k])), shape=(U, I)).astype(np.intc) for k in range(T)]
    trn = sp.csr_matrix((vals, (rows, cols)), shape=(U, I)).astype(np.intc)
    if stored_zero:      # explicit zeros in the stored data: `temLabel != 0` / `temLabel[item] == 0` must see through them
        for m in sub + [trn]:
            m.data[::7] = 0
    return seqs, tst, sub, trn


def main():
    shim, model, NNs = load_reference()
    args = model.args
    cases = {
        # name: (U, I, T, sslNum, train_sample_num, pred_num, pos_length, args.batch, batch sizes, stored zeros)
        "a": (60, 90, 3, 20, 40, 5, 200, 32, [32, 28], False),
        "b_short_pos": (45, 30, 4, 3, 7, 2, 6, 16, [16, 16, 13], False),
        "c_stored_zeros": (40, 64, 2, 5, 10, 5, 12, 20, [20, 20], True),
    }
    for name, (U, I, T, ssl, tsn, pred_num, pos_len, abatch, batches, zeros) in cases.items():
        rng = np.random.default_rng({"a": 1, "b_short_pos": 2, "c_stored_zeros": 3}[name])
        seqs, tst, sub, trn = make_data(rng, U, I, T, stored_zero=zeros)
        args.user, args.item, args.graphNum, args.sslNum = U, I, T, ssl
        args.pred_num, args.pos_length, args.batch = pred_num, pos_len, abatch
        rec = _Obj()
        rec.handler = _Obj()
        rec.handler.sequence, rec.handler.tstInt, rec.handler.item_with_pop = seqs, np.array(tst, dtype=object), None
        np.random.seed(100); random.seed(100)                      # main.py:21-22
        out = dict(U=U, I=I, T=T, sslNum=ssl, train_sample_num=tsn, pred_num=pred_num, pos_length=pos_len, batch=abatch,
                   n_batches=len(batches), tst_int=np.array([-1 if x is None else x for x in tst], np.int32),
                   seq_ptr=np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64),
                   seq_items=np.concatenate(seqs).astype(np.int32))
        for k, m in enumerate(sub + [trn]):
            tag = "trn" if k == T else "sub%d" % k
            out[tag + "_indptr"], out[tag + "_indices"], out[tag + "_data"] = m.indptr, m.indices, m.data
        sf = np.random.permutation(U)                              # model.py:342
        out["perm"] = sf
        st = 0
        for b, bs in enumerate(batches):
            bat = sf[st:st + bs]; st += bs
            out["bat%d" % b] = bat.astype(np.int32)
            uL, iL, seq, mask, uLs = model.Recommender.sampleTrainBatch(rec, bat, trn, None, tsn)      # model.py:352
            out["trn_uLocs%d" % b], out["trn_iLocs%d" % b], out["trn_uLocs_seq%d" % b] = map(np.asarray, (uL, iL, uLs))
            out["trn_sequence%d" % b], out["trn_mask%d" % b] = np.asarray(seq), np.asarray(mask)
            assert np.asarray(seq).shape == (abatch, pos_len)
            suL, siL, suLs = model.Recommender.sampleSslBatch(rec, bat, sub, False)                      # model.py:353
            for k in range(T):
                out["ssl_uLocs%d_%d" % (b, k)], out["ssl_iLocs%d_%d" % (b, k)], out["ssl_uLocs_seq%d_%d" % (b, k)] = \
                    np.asarray(suL[k], np.int64), np.asarray(siL[k], np.int64), np.asarray(suLs[k], np.int64)
            npst, pyst = np.random.get_state(), random.getstate()
            out["np_key%d" % b], out["np_pos%d" % b] = npst[1].copy(), npst[2]
            out["py_key%d" % b] = np.array(pyst[1], dtype=np.uint32)
        np.savez_compressed(os.path.join(HERE, "sampler_%s.npz" % name), **out)
        print("sampler", name, "batches", batches, "train samples", [len(out["trn_uLocs%d" % b]) for b in range(len(batches))],
              "ssl samples b0", [len(out["ssl_uLocs0_%d" % k]) for k in range(T)])
    # what the reference does with a user of fewer than 3 interactions
    args.user, args.item, args.pos_length, args.batch, args.pred_num = 2, 9, 5, 2, 5
    rec = _Obj(); rec.handler = _Obj()
    rec.handler.sequence, rec.handler.tstInt, rec.handler.item_with_pop = [[1, 2], [3, 4, 5, 6]], np.array([None, None], dtype=object), None
    trn = sp.csr_matrix(([1, 1, 1, 1, 1, 1], ([0, 0, 1, 1, 1, 1], [1, 2, 3, 4, 5, 6])), shape=(2, 9))
    try:
        model.Recommender.sampleTrainBatch(rec, np.array([0, 1]), trn, None, 3)
        err = "ran"
    except Exception as e:   # noqa: BLE001
        err = "%s: %s" % (type(e).__name__, e)
    print("reference on a 2-interaction user ->", err)
    np.savez_compressed(os.path.join(HERE, "sampler_errors.npz"), short_sequence=np.array(err))


if __name__ == "__main__":
    main()
