"""Throughput of the parity-mode samplers (host C, sagnn_b200.ReferenceStream) next to the reference's own Python
samplers on a Gowalla-shaped data set.  BUILD CONTAINER ONLY (imports /root/reference/model.py over the TF stand-in).

    python scripts/sampler_speed.py [--batch 512] [--batches 4]
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--batches", type=int, default=4)
    a = ap.parse_args()
    from make_golden_model import load_reference
    shim, model, NNs = load_reference()
    sys.argv = ["x"]
    from sagnn_b200 import data_handler as dh
    from sagnn_b200.np_sampler import ReferenceStream
    import scipy.sparse as sp
    g = dh.make_named("gowalla", seed=100)
    U, I, T = g.n_user, g.n_item, g.graph_num
    trn = sp.csr_matrix(sum((m != 0).astype(np.intc) for m in g.sub_mat))
    trn.sum_duplicates(); trn.sort_indices()
    rng = np.random.default_rng(0)
    seqs = []
    for u in range(U):                                   # a user's items in a random "time" order; at least 3
        items = trn.indices[trn.indptr[u]:trn.indptr[u + 1]]
        if len(items) < 3:
            items = np.unique(np.concatenate([items, rng.integers(0, I, size=3)]))
        seqs.append([int(x) for x in rng.permutation(items)])
    tst = [None if rng.random() < 0.3 else int(rng.integers(0, I)) for _ in range(U)]
    args = model.args
    args.user, args.item, args.graphNum, args.sslNum, args.pred_num, args.pos_length, args.batch = U, I, T, 20, 5, 200, a.batch

    class O: pass
    rec = O(); rec.handler = O()
    rec.handler.sequence, rec.handler.tstInt, rec.handler.item_with_pop = seqs, np.array(tst, dtype=object), None
    np.random.seed(100); random.seed(100)
    rs = ReferenceStream(100, 100)
    perm = np.random.permutation(U); assert np.array_equal(perm, rs.np_permutation(U))
    t_ref = t_ours = 0.0
    ours_per_batch = []
    n_samples = 0
    for b in range(a.batches):
        bat = perm[b * a.batch:(b + 1) * a.batch]
        t0 = time.perf_counter()
        r1 = model.Recommender.sampleTrainBatch(rec, bat, trn, None, 40)
        r2 = model.Recommender.sampleSslBatch(rec, bat, g.sub_mat, False)
        t_ref += time.perf_counter() - t0
        t0 = time.perf_counter()
        o1 = rs.sample_train_batch(bat, trn, seqs, tst, 40, pred_num=5, pos_length=200, batch_pad=a.batch)
        o2 = rs.sample_ssl_batch(bat, g.sub_mat, 20)
        ours_per_batch.append(time.perf_counter() - t0)
        t_ours += ours_per_batch[-1]
        assert np.array_equal(o1[0], np.asarray(r1[0])) and np.array_equal(o1[1], np.asarray(r1[1]))
        assert np.array_equal(o1[2], np.asarray(r1[2])) and np.array_equal(o1[3], np.asarray(r1[3]))
        for k in range(T):
            assert np.array_equal(o2[1][k], np.asarray(r2[1][k], np.int64))
        n_samples += len(o1[0]) + sum(len(x) for x in o2[0])
    st = np.random.get_state()
    assert np.array_equal(rs.to_numpy()[1], st[1]) and rs.to_numpy()[2] == st[2] and rs.to_python() == random.getstate()
    steady = float(np.mean(ours_per_batch[1:])) if len(ours_per_batch) > 1 else ours_per_batch[0]
    print("gowalla-shaped U=%d I=%d T=%d, batch %d x %d: reference %.1f ms/batch; parity mode %.2f ms/batch after the first "
          "(first: %.1f ms, converts the static inputs once) = %.0fx; %d samples/batch, outputs and generator states identical"
          % (U, I, T, a.batch, a.batches, 1e3 * t_ref / a.batches, 1e3 * steady, 1e3 * ours_per_batch[0],
             (t_ref / a.batches) / steady, n_samples // a.batches))


if __name__ == "__main__":
    main()
