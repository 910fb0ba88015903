"""Stream-exact samplers (parity mode; host code of the C ABI, no GPU needed): the MT19937 primitives
against numpy's legacy RandomState and CPython's ``random`` themselves, and ``sampleSslBatch`` / ``sampleTrainBatch``
/ ``negSamp`` BIT FOR BIT against fixtures produced by calling the reference's own functions under main.py's seeds
(tests/golden/make_golden_sampler.py; model.py:252-339, DataHandler.py:28-41)."""
import glob
import os
import random

import numpy as np
import pytest
import scipy.sparse as sp

from sagnn_b200.np_sampler import ReferenceStream
from sagnn_b200._lib import SagnnError

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(p for p in glob.glob(os.path.join(GOLD, "sampler_*.npz")) if not p.endswith("sampler_errors.npz"))


@pytest.mark.parametrize("seed", [0, 1, 100, 20261019, 2**32 - 1])
def test_primitives_follow_numpy_and_random_streams(seed):
    rs = ReferenceStream(seed, seed)
    np.random.seed(seed); random.seed(seed)
    for n in (1, 2, 3, 7, 100, 52621, 2**20, 2**20 + 1, 2**31, 2**32 - 1, 2**32, 2**32 + 5, 2**40 + 3):
        assert np.array_equal(np.random.randint(0, n, size=33), rs.np_randint(0, n, 33))
        assert np.random.choice(n) == rs.np_randint(0, n)          # np.random.choice(int) = randint(0, int)
        assert [random.randint(1, n) for _ in range(9)] == [rs.py_randint(1, n) for _ in range(9)]
    assert np.array_equal(np.random.randint(-5, 17, size=20), rs.np_randint(-5, 17, 20))
    for n in (0, 1, 2, 5, 1000, 48653):
        assert np.array_equal(np.random.permutation(n), rs.np_permutation(n))
    arr = np.arange(10, 400, 3)
    assert np.array_equal(np.random.choice(arr, 40), arr[rs.np_randint(0, len(arr), 40)])
    st, want = rs.to_numpy(), np.random.get_state()
    assert np.array_equal(st[1], want[1]) and st[2] == want[2]
    assert rs.to_python() == random.getstate()


def test_wide_python_seeds_and_state_takeover():
    for seed in (2**32 + 7, 2**63 + 11):
        rs = ReferenceStream(1, seed); random.seed(seed)
        assert [random.randint(0, 1000) for _ in range(5)] == [rs.py_randint(0, 1000) for _ in range(5)]
    np.random.seed(7); np.random.rand(3); random.seed(9); random.random()     # streams somewhere in the middle
    rs = ReferenceStream().from_numpy().from_python()
    assert np.array_equal(np.random.randint(0, 1000, 50), rs.np_randint(0, 1000, 50))
    assert random.randint(0, 99) == rs.py_randint(0, 99)
    np.random.set_state(rs.to_numpy()); random.setstate(rs.to_python())         # and back
    assert np.random.randint(0, 10**6) == rs.np_randint(0, 10**6) and random.randint(0, 10**6) == rs.py_randint(0, 10**6)
    with pytest.raises(SagnnError):
        rs.np_randint(5, 5)
    with pytest.raises(SagnnError):
        rs.py_randint(3, 2)


def _mats(fx):
    T, U, I = int(fx["T"]), int(fx["U"]), int(fx["I"])
    csr = lambda tag: sp.csr_matrix((fx[tag + "_data"], fx[tag + "_indices"], fx[tag + "_indptr"]), shape=(U, I))
    seqs = [fx["seq_items"][fx["seq_ptr"][u]:fx["seq_ptr"][u + 1]].tolist() for u in range(U)]
    tst = [None if x < 0 else int(x) for x in fx["tst_int"]]
    return [csr("sub%d" % k) for k in range(T)], csr("trn"), seqs, tst


def test_sampler_fixtures_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-4] for p in CASES])
def test_samplers_reproduce_the_reference_bit_for_bit(path):
    """An epoch prefix as trainEpoch runs it (model.py:342-356): permutation, then per batch sampleTrainBatch and
    sampleSslBatch -- outputs AND the generator states after every batch equal the reference's."""
    fx = np.load(path)
    sub, trn, seqs, tst = _mats(fx)
    T = int(fx["T"])
    rs = ReferenceStream(100, 100)                                  # main.py:21-22
    sf = rs.np_permutation(int(fx["U"]))
    assert np.array_equal(sf, fx["perm"])
    for b in range(int(fx["n_batches"])):
        bat = fx["bat%d" % b]
        uL, iL, seq, mask, uLs = rs.sample_train_batch(bat, trn, seqs, tst, int(fx["train_sample_num"]), pred_num=int(fx["pred_num"]),
                                                       pos_length=int(fx["pos_length"]), batch_pad=int(fx["batch"]))
        assert np.array_equal(uL, fx["trn_uLocs%d" % b]) and np.array_equal(iL, fx["trn_iLocs%d" % b])
        assert np.array_equal(uLs, fx["trn_uLocs_seq%d" % b])
        assert np.array_equal(seq, fx["trn_sequence%d" % b]) and np.array_equal(mask, fx["trn_mask%d" % b])
        suL, siL, suLs = rs.sample_ssl_batch(bat, sub, int(fx["sslNum"]))
        for k in range(T):
            assert np.array_equal(suL[k], fx["ssl_uLocs%d_%d" % (b, k)]), (b, k)
            assert np.array_equal(siL[k], fx["ssl_iLocs%d_%d" % (b, k)]), (b, k)
            assert np.array_equal(suLs[k], fx["ssl_uLocs_seq%d_%d" % (b, k)]), (b, k)
        st = rs.to_numpy()
        assert np.array_equal(st[1], fx["np_key%d" % b]) and st[2] == int(fx["np_pos%d" % b])
        assert np.array_equal(np.array(rs.to_python()[1], dtype=np.uint32), fx["py_key%d" % b])


def test_sampler_edge_cases():
    """Users without enough items in an interval draw (and discard) one item like model.py:321-323; a user with
    fewer than 3 interactions makes the reference raise (recorded in sampler_errors.npz) -- so does the C ABI."""
    U, I = 4, 11
    sub = [sp.csr_matrix(([5, 6, 7], ([0, 0, 2], [1, 9, 4])), shape=(U, I)).astype(np.intc)]
    rs = ReferenceStream(3, 3); np.random.seed(3)
    uL, iL, uLs = rs.sample_ssl_batch(np.array([1, 0, 2, 3]), sub, ssl_num=2)
    # reference order of draws: user 1 (empty) -> choice(I); user 0 (2 items) -> choice(posset, 2); users 2, 3 -> choice(I)
    np.random.choice(I)
    want = np.array([1, 9])[np.random.randint(0, 2, size=2)]
    np.random.choice(I); np.random.choice(I)
    assert iL[0].tolist() == want.tolist() and uL[0].tolist() == [0, 0] and uLs[0].tolist() == [1, 1]
    assert np.array_equal(rs.to_numpy()[1], np.random.get_state()[1]) and rs.to_numpy()[2] == np.random.get_state()[2]
    empty = rs.sample_ssl_batch(np.zeros(0, np.int32), sub, ssl_num=2)
    assert [len(x[0]) for x in empty] == [0, 0, 0]
    assert "ValueError" in str(np.load(os.path.join(GOLD, "sampler_errors.npz"))["short_sequence"])
    trn = sp.csr_matrix(([1] * 6, ([0, 0, 1, 1, 1, 1], [1, 2, 3, 4, 5, 6])), shape=(2, 9))
    with pytest.raises(SagnnError, match="fewer than 3"):
        rs.sample_train_batch(np.array([0, 1]), trn, [[1, 2], [3, 4, 5, 6]], [None, None], 3, pos_length=5)
    with pytest.raises(SagnnError):
        rs.sample_ssl_batch(np.array([U]), sub, ssl_num=2)          # user id out of range


# ---- the samplers' numpy restatement (oracle/sampler_oracle.py): pinned by the same fixtures, then used as the
#      live checker for many random cases ---------------------------------------------------------------------------
from oracle import sampler_oracle as so


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[8:-4] for p in CASES])
def test_sampler_oracle_reproduces_the_reference_fixtures(path):
    fx = np.load(path)
    sub, trn, seqs, tst = _mats(fx)
    T = int(fx["T"])
    np_rng, py_rng = np.random.RandomState(100), random.Random(100)
    assert np.array_equal(np_rng.permutation(int(fx["U"])), fx["perm"])
    for b in range(int(fx["n_batches"])):
        bat = fx["bat%d" % b]
        uL, iL, seq, mask, uLs = so.sample_train_batch(np_rng, py_rng, bat, trn, seqs, tst, int(fx["train_sample_num"]),
                                                       int(fx["pred_num"]), int(fx["pos_length"]), int(fx["batch"]), int(fx["I"]))
        assert np.array_equal(uL, fx["trn_uLocs%d" % b]) and np.array_equal(iL, fx["trn_iLocs%d" % b])
        assert np.array_equal(uLs, fx["trn_uLocs_seq%d" % b])
        assert np.array_equal(seq, fx["trn_sequence%d" % b]) and np.array_equal(mask, fx["trn_mask%d" % b])
        suL, siL, suLs = so.sample_ssl_batch(np_rng, bat, sub, int(fx["sslNum"]), int(fx["I"]))
        for k in range(T):
            assert np.array_equal(suL[k], fx["ssl_uLocs%d_%d" % (b, k)]) and np.array_equal(siL[k], fx["ssl_iLocs%d_%d" % (b, k)])
            assert np.array_equal(suLs[k], fx["ssl_uLocs_seq%d_%d" % (b, k)])
        st = np_rng.get_state()
        assert np.array_equal(st[1], fx["np_key%d" % b]) and st[2] == int(fx["np_pos%d" % b])
        assert np.array_equal(np.array(py_rng.getstate()[1], dtype=np.uint32), fx["py_key%d" % b])


@pytest.mark.parametrize("seed", range(12))
def test_samplers_equal_the_pinned_oracle_on_random_cases(seed):
    """Random shapes, densities, batch compositions (repeated users, users with no items in an interval, dense
    users for whom negSamp rejects most draws, stored zeros), seeds: C ABI == oracle, draws and stream positions."""
    rng = np.random.default_rng(1000 + seed)
    U, I, T = int(rng.integers(5, 80)), int(rng.integers(12, 120)), int(rng.integers(1, 6))
    seqs, tst, rows, cols, ks = [], [], [], [], []
    for u in range(U):
        n = int(rng.integers(3, max(4, min(I - 3, 3 + int(rng.integers(0, I))))))      # up to nearly every item
        items = rng.choice(I, size=n, replace=False)
        seqs.append([int(x) for x in items])
        tst.append(None if rng.random() < 0.4 else int(rng.integers(0, I)))
        rows += [u] * n; cols += [int(x) for x in items]; ks += [int(x) for x in rng.integers(0, T, size=n)]
    rows, cols, ks = np.asarray(rows), np.asarray(cols), np.asarray(ks)
    vals = rng.integers(1, 1000, size=len(rows)).astype(np.intc)
    if seed % 3 == 0:
        vals[rng.random(len(vals)) < 0.2] = 0                                             # stored explicit zeros
    sub = [sp.csr_matrix((vals[ks == k], (rows[ks == k], cols[ks == k])), shape=(U, I)) for k in range(T)]
    trn = sp.csr_matrix((vals, (rows, cols)), shape=(U, I))
    ssl_num, tsn, pred_num, pos_len = int(rng.integers(0, 8)), int(rng.integers(0, 9)), int(rng.integers(0, 7)), int(rng.integers(1, 20))
    s1, s2 = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32))
    rs = ReferenceStream(s1, s2)
    np_rng, py_rng = np.random.RandomState(s1), random.Random(s2)
    for _ in range(3):
        bat = rng.integers(0, U, size=int(rng.integers(1, 25))).astype(np.int32)          # users may repeat
        pad = len(bat) + int(rng.integers(0, 4))
        got = rs.sample_train_batch(bat, trn, seqs, tst, tsn, pred_num=pred_num, pos_length=pos_len, batch_pad=pad)
        want = so.sample_train_batch(np_rng, py_rng, bat, trn, seqs, tst, tsn, pred_num, pos_len, pad, I)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        got = rs.sample_ssl_batch(bat, sub, ssl_num)
        want = so.sample_ssl_batch(np_rng, bat, sub, ssl_num, I)
        for gl, wl in zip(got, want):
            for g, w in zip(gl, wl):
                assert np.array_equal(g, w)
        st = np_rng.get_state()
        assert np.array_equal(rs.to_numpy()[1], st[1]) and rs.to_numpy()[2] == st[2]
        assert rs.to_python() == py_rng.getstate()
