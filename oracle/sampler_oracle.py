"""numpy / ``random`` restatement of the reference's host samplers (TEST INFRASTRUCTURE ONLY): what
``Recommender.sampleSslBatch`` (LIU-YUXI/SA-GNN model.py:304-339), ``Recommender.sampleTrainBatch``
(model.py:252-302) and ``negSamp`` (DataHandler.py:28-41) draw and emit, written against explicit generator
objects (``np.random.RandomState`` = numpy's global stream, ``random.Random`` = the ``random`` module) so that tests
can run many seeded cases side by side with the C ABI's parity-mode samplers.

PINNED: ``tests/test_np_stream.py`` checks these functions against ``tests/golden/sampler_*.npz``, which were produced
by calling the reference's own functions under main.py's seeds (tests/golden/make_golden_sampler.py).
"""
from __future__ import annotations

import numpy as np


def sample_ssl_batch(np_rng, bat_ids, sub_mats, ssl_num, n_item):
    """model.py:304-339.  Per interval k and batch position b: posset = columns with a non-zero value in row
    bat_ids[b] of sub_mats[k]; s = min(ssl_num, |posset| // 2); s == 0 draws (and drops) one item id; otherwise
    2s draws with replacement from posset, emitted as s interleaved (first-half, second-half) pairs."""
    u_out, i_out, s_out = [], [], []
    for m in sub_mats:
        dense = m[bat_ids].toarray() if len(bat_ids) else np.zeros((0, m.shape[1]))
        u, i, s = [], [], []
        for b, uid in enumerate(bat_ids):
            posset = np.flatnonzero(dense[b] != 0)
            n = min(ssl_num, len(posset) // 2)
            if n == 0:
                np_rng.choice(n_item)
                continue
            drawn = np_rng.choice(posset, 2 * n)
            for j in range(n):
                u += [uid, uid]; s += [b, b]; i += [drawn[j], drawn[n + j]]
        u_out.append(np.asarray(u, np.int64)); i_out.append(np.asarray(i, np.int64)); s_out.append(np.asarray(s, np.int64))
    return u_out, i_out, s_out


def neg_samp(np_rng, label_row, n, n_item, excluded):
    """DataHandler.py:28-41: uniform item ids until n of them have no label and are not in ``excluded``."""
    out = []
    while len(out) < n:
        r = np_rng.choice(n_item)
        if label_row[r] == 0 and r not in excluded:
            out.append(r)
    return out


def sample_train_batch(np_rng, py_rng, bat_ids, label_mat, sequences, tst_int, train_sample_num, pred_num, pos_length,
                       batch_pad, n_item):
    """model.py:252-302.  Returns (uLocs, iLocs, sequence [batch_pad, pos_length] int64, mask float64, uLocs_seq):
    positives of all users in batch order, then their negatives."""
    dense = label_mat[bat_ids].toarray()
    pos_u, pos_i, pos_s, neg_i = [], [], [], []
    seq = np.zeros((batch_pad, pos_length), np.int64)
    mask = np.zeros((batch_pad, pos_length))
    for b, uid in enumerate(bat_ids):
        full = list(sequences[uid])
        posset = full[:-1]
        n = min(train_sample_num, len(posset))
        choose = 1
        if n == 0:
            np_rng.choice(n_item)
        else:
            choose = py_rng.randint(1, max(min(pred_num + 1, len(posset) - 3), 1))
            negs = neg_samp(np_rng, dense[b], n, n_item, [full[-1], tst_int[uid]])
            pos_u += [uid] * n; pos_s += [b] * n; pos_i += [posset[-choose]] * n; neg_i += negs
        hist = posset[:-choose]
        if not hist:
            raise ValueError("could not broadcast input array from shape (0,) into shape (%d,)" % pos_length)
        if len(hist) <= pos_length:
            seq[b, -len(hist):] = hist; mask[b, -len(hist):] = 1
        else:
            seq[b] = hist[-pos_length:]; mask[b] = 1
    arr = lambda x: np.asarray(x, np.int64)
    return arr(pos_u + pos_u), arr(pos_i + neg_i), seq, mask, arr(pos_s + pos_s)
