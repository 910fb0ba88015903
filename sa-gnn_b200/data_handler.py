"""Host-side mirror of the reference's data layer for the propagation path.

Mirrors (names, argument meaning, return layout, edge cases) the three pieces of
``DataHandler.py`` the interval-graph propagation consumes:

  * ``transpose``        -- DataHandler.py:9-11
  * ``transToLsts``      -- DataHandler.py:47-69 (adjacency list, dead int32
                            "normalisation", empty-matrix fallback edge (0,0))
  * ``LoadData``'s ``trn_mat_time`` part -- DataHandler.py:92-94,126-129

plus a writer of the same pickle layout and the synthetic power-law interval
graph generator of SURVEY section 8(d) / appendix D (the real ``trn_mat_time``
files are not shipped with the reference).

Pure numpy/scipy host code: it only *prepares* inputs.  The CSR/CSC
construction the model actually runs on is done on the device by the plan
(``propagate.build_plan``); nothing here is a CPU fallback for it.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

TS_LO, TS_HI = 1388534400, 1406073600   # timestamp range of the stored values


# --------------------------------------------------------------------------
# reference-named helpers
# --------------------------------------------------------------------------
def transpose(mat):
    """DataHandler.py:9-11 -- CSR of ``mat``'s transpose (rows = items, user ids ascending)."""
    return sp.csr_matrix(sp.coo_matrix(mat).transpose())


def transToLsts(mat, mask=False, norm=False):
    """DataHandler.py:47-69 -- ``(indices int32 [E,2], data int32 [E], [R, C])``.

    Vectorised, same results: COO order of the canonical CSR; ``norm`` multiplies
    every stored value by ``rowD[row]*colD[col]`` in float64 and stores it back
    into the int32 array (truncation -- all zeros for real data);
    ``mask`` applies the reference's random half mask (:62-64); an empty matrix
    becomes the single edge ``[[0, 0]]`` with data ``[0]`` (:66-68).
    """
    shape = [int(mat.shape[0]), int(mat.shape[1])]
    coo = sp.coo_matrix(mat)
    row = np.asarray(coo.row, dtype=np.int32)
    col = np.asarray(coo.col, dtype=np.int32)
    indices = np.stack([row, col], axis=1) if row.size else np.zeros((0, 2), np.int32)
    data = np.asarray(coo.data).astype(np.int32)
    if norm and data.size:
        row_d = 1.0 / (np.sqrt(np.asarray(mat.sum(axis=1)).reshape(-1) + 1e-8) + 1e-8)
        col_d = 1.0 / (np.sqrt(np.asarray(mat.sum(axis=0)).reshape(-1) + 1e-8) + 1e-8)
        data = ((data * row_d[row]) * col_d[col]).astype(np.int32)
    if mask:
        data = data * ((np.random.uniform(size=data.shape) > 0.5) * 1.0)
    if indices.shape[0] == 0:
        indices = np.array([[0, 0]], dtype=np.int32)
        data = np.array([0], dtype=np.int32)
    return indices, data, shape


trans_to_lsts = transToLsts


# --------------------------------------------------------------------------
# trn_mat_time  
# --------------------------------------------------------------------------
@dataclass
class IntervalGraphs:
    """What ``DataHandler.LoadData`` keeps for the path: ``subMat`` (T CSR, intc
    timestamps), the global matrix (only its shape is used) and ``timeMat``."""
    n_user: int
    n_item: int
    sub_mat: list                      # T x csr_matrix U x I, dtype intc  (trnMat[1])
    trn_mat: object = None             # csr U x I float64                 (trnMat[0])
    time_mat: object = None            # csr U x I intc                    (trnMat[2])
    meta: dict = field(default_factory=dict)

    @property
    def graph_num(self):
        return len(self.sub_mat)

    @property
    def nnz(self):
        return [int(m.nnz) for m in self.sub_mat]


def load_trn_mat_time(path, graph_num=None):
    """DataHandler.py:92-94,126-129: unpickle ``[trnMat, subMat[T], timeMat]``;
    ``args.user, args.item = trnMat[0].shape``; ``--graphNum`` may select a prefix
    of the stored intervals (model.py:230-231)."""
    with open(path, "rb") as fs:
        trn = pickle.load(fs)
    n_user, n_item = trn[0].shape
    sub = list(trn[1])
    if graph_num is not None:
        if graph_num > len(sub):
            raise IndexError("graphNum %d > %d stored intervals" % (graph_num, len(sub)))
        sub = sub[:graph_num]
    return IntervalGraphs(int(n_user), int(n_item), [sp.csr_matrix(m) for m in sub], trn[0], trn[2])


def write_trn_mat_time(path, graphs: IntervalGraphs):
    """Writes the pickle layout of ``preprocess_to_trnmat.ipynb:1896``."""
    U, I = graphs.n_user, graphs.n_item
    trn = graphs.trn_mat
    if trn is None:      # interaction counts: every pair lives in exactly one interval
        trn = sp.csr_matrix((U, I), dtype=np.float64)
        for m in graphs.sub_mat:
            trn = trn + (m != 0).astype(np.float64)
        trn = sp.csr_matrix(trn)
    tm = graphs.time_mat
    if tm is None:       # last interval id per pair; interval 0 vanishes (DOK drops zeros)
        tm = sp.csr_matrix((U, I), dtype=np.intc)
        for k, m in enumerate(graphs.sub_mat):
            if k:
                tm = tm + ((m != 0).astype(np.intc) * k)
        tm = sp.csr_matrix(tm).astype(np.intc)
    with open(path, "wb") as fs:
        pickle.dump([trn, [m.astype(np.intc) for m in graphs.sub_mat], tm], fs)


# --------------------------------------------------------------------------
# producer of trn_mat_time (preprocess_to_trnmat.ipynb cells 7, 13, 14): vectorised host
# mirror of `trans` / `trans_sub`, pinned bit-exactly by tests/golden/trnmat_*.npz (made by exec'ing
# the notebook's own cells).  The notebook walks list[user] -> {item: [timestamps]} in Python loops;
# here the same events arrive as flat (user, item, timestamp) arrays in that iteration order.
# --------------------------------------------------------------------------
TS_MINN_INIT, TS_MAXX_INIT = 1647180684, 0      # the notebook's initial minn / maxx (cell 13)


def interaction_to_triples(interaction):
    """Flattens the notebook's ``trnInt`` (``list[U]`` of ``None | {item: [timestamps] | None}``) into
    ``(users, items, times)`` int64 arrays in the order `trans` / `trans_sub` visit the events."""
    us, its, ts = [], [], []
    for usr, data in enumerate(interaction):
        if data is None:
            continue
        for col in data:
            if data[col] is not None:
                for one in data[col]:
                    us.append(usr)
                    its.append(col)
                    ts.append(one)
    return np.array(us, np.int64), np.array(its, np.int64), np.array(ts, np.int64)


def trans(users, items, times, n_user, n_item):
    """`trans` (notebook cell 13): the global ``U x I`` float64 interaction-count matrix ``trnMat[0]``
    plus the running ``(minn, maxx)`` timestamps that `trans_sub` buckets with."""
    u, i, t = (np.asarray(x, np.int64) for x in (users, items, times))
    minn = min(TS_MINN_INIT, int(t.min())) if t.size else TS_MINN_INIT
    maxx = max(TS_MAXX_INIT, int(t.max())) if t.size else TS_MAXX_INIT
    mat = sp.csr_matrix((np.ones(t.size, np.float64), (u, i)), shape=(n_user, n_item))   # duplicates are summed
    return mat, minn, maxx


def trans_sub(users, items, times, n_user, n_item, graph_num, minn, maxx):
    """`trans_sub` (notebook cell 7): splits the events into ``graph_num`` equal time intervals,
    ``interval id = int((t - minn) / ((maxx - minn) / graph_num))`` clamped to ``graph_num - 1``; interval
    graph k holds, per ``(user, item)``, the timestamp of the pair's FIRST event (in visiting order)
    that falls into interval k.  ``timeMat[user, item]`` = the interval id of the pair's last such
    first-event, and -- the notebook fills a DOK matrix -- pairs whose value is 0 are not stored.
    Returns ``(list of graph_num csr U x I intc, timeMat csr intc)``; ``trn_mat_time`` is
    ``[trans(...)[0], sub_mats, time_mat]`` (cell 14)."""
    u, i, t = (np.asarray(x, np.int64) for x in (users, items, times))
    if maxx == minn:
        raise ZeroDivisionError("all events share one timestamp: the notebook's interval width is 0")
    interval = (maxx - minn) / graph_num                       # Python float, as in the notebook
    g = np.minimum(((t - minn) / interval).astype(np.int64), graph_num - 1)
    # first event of every (interval, user, item): np.unique returns first occurrences
    key = (g * n_user + u) * n_item + i
    _, first = np.unique(key, return_index=True)
    first.sort()                                               # back to visiting order
    gu, uu, ii, tt = g[first], u[first], i[first], t[first]
    subs = []
    for k in range(graph_num):
        m = gu == k
        subs.append(sp.csr_matrix((tt[m].astype(np.intc), (uu[m], ii[m])), shape=(n_user, n_item), dtype=np.intc))
    # timeMat: the interval of the LAST appended event of each pair (later assignments overwrite)
    pair = uu * n_item + ii
    order = np.lexsort((first, pair))
    last = np.ones(order.size, bool)
    last[:-1] = pair[order][1:] != pair[order][:-1]
    sel = order[last]
    keep = gu[sel] != 0                                         # DOK drops explicit zeros
    tm = sp.csr_matrix((gu[sel][keep].astype(np.intc), (uu[sel][keep], ii[sel][keep])),
                       shape=(n_user, n_item), dtype=np.intc)
    return subs, tm


def make_trn_mat_time(users, items, times, n_user, n_item, graph_num):
    """Cells 13-14 end to end: ``IntervalGraphs`` ready for `write_trn_mat_time` / `write_trn_mat_bin`."""
    trn, minn, maxx = trans(users, items, times, n_user, n_item)
    subs, tm = trans_sub(users, items, times, n_user, n_item, graph_num, minn, maxx)
    return IntervalGraphs(int(n_user), int(n_item), subs, trn, tm, meta={"minn": minn, "maxx": maxx})


# --------------------------------------------------------------------------
# binary CSR container: the same content as trn_mat_time[1] (+ shape), but as raw
# little-endian arrays that can be memory-mapped -- no unpickling of T scipy objects, O(1) open,
# intervals can be read (or sharded over ranks) independently.  Layout, all offsets 64-byte aligned:
#   0   8s  magic "SAGNNCSR"      8  u32 version (1)     12 u32 T      16 u64 U      24 u64 I
#   32  u32 flags (bit 0: stored values present)          36 u32 reserved        40 u64 nnz[T]
#   then per interval k: indptr int64 [U+1] | indices int32 [nnz_k] | data int32 [nnz_k] (if flag)
# indptr / indices are the canonical CSR of subMat[k] (sorted column ids, no duplicates), data the
# intc timestamps (preprocess_to_trnmat.ipynb:379) -- exactly what transToLsts consumes.
# --------------------------------------------------------------------------
_BIN_MAGIC = b"SAGNNCSR"
_BIN_VERSION = 1


def _align64(n):
    return (n + 63) & ~63


def write_trn_mat_bin(path, graphs, with_values=True):
    """Writes ``graphs`` (IntervalGraphs or a list of U x I scipy matrices) as a binary CSR container."""
    mats = graphs.sub_mat if isinstance(graphs, IntervalGraphs) else list(graphs)
    if not mats:
        raise ValueError("need at least one interval")
    mats = [sp.csr_matrix(m) for m in mats]
    for m in mats:
        m.sum_duplicates()
        m.sort_indices()
    U, I = mats[0].shape
    if any(m.shape != (U, I) for m in mats):
        raise ValueError("all intervals must have the same shape")
    T = len(mats)
    with open(path, "wb") as f:
        head = bytearray(_align64(40 + 8 * T))
        head[0:8] = _BIN_MAGIC
        head[8:40] = np.array([_BIN_VERSION, T], "<u4").tobytes() + np.array([U, I], "<u8").tobytes() + \
            np.array([1 if with_values else 0, 0], "<u4").tobytes()
        head[40:40 + 8 * T] = np.array([m.nnz for m in mats], "<u8").tobytes()
        f.write(head)
        for m in mats:
            parts = [m.indptr.astype("<i8"), m.indices.astype("<i4")]
            if with_values:
                parts.append(m.data.astype("<i4"))
            for a in parts:
                b = a.tobytes()
                f.write(b)
                f.write(b"\0" * (_align64(len(b)) - len(b)))


def load_trn_mat_bin(path, graph_num=None, mmap=True):
    """Opens a binary CSR container: returns ``IntervalGraphs`` whose ``sub_mat`` are scipy CSR
    matrices backed by the memory-mapped file (``mmap=False``: read into memory).  ``graph_num``
    selects a prefix of the stored intervals like ``--graphNum`` (model.py:230-231)."""
    raw = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    if raw.size < 40 or bytes(raw[0:8]) != _BIN_MAGIC:
        raise ValueError("%s is not a SAGNNCSR container" % path)
    version, T = (int(x) for x in raw[8:16].view("<u4"))
    if version != _BIN_VERSION:
        raise ValueError("unsupported SAGNNCSR version %d" % version)
    U, I = (int(x) for x in raw[16:32].view("<u8"))
    flags = int(raw[32:36].view("<u4")[0])
    nnz = [int(x) for x in raw[40:40 + 8 * T].view("<u8")]
    if graph_num is not None:
        if graph_num > T:
            raise IndexError("graphNum %d > %d stored intervals" % (graph_num, T))
        T = graph_num
    off = _align64(40 + 8 * len(nnz))
    mats = []
    for k in range(len(nnz)):
        sizes = [8 * (U + 1), 4 * nnz[k]] + ([4 * nnz[k]] if flags & 1 else [])
        if off + sum(_align64(x) for x in sizes) > raw.size:
            raise ValueError("%s is truncated (interval %d)" % (path, k))
        if k < T:
            indptr = raw[off:off + sizes[0]].view("<i8")
            indices = raw[off + _align64(sizes[0]):off + _align64(sizes[0]) + sizes[1]].view("<i4")
            if flags & 1:
                o2 = off + _align64(sizes[0]) + _align64(sizes[1])
                data = raw[o2:o2 + sizes[2]].view("<i4")
            else:
                data = np.ones(nnz[k], dtype=np.intc)
            if int(indptr[-1]) != nnz[k] or (nnz[k] and (int(indices.min()) < 0 or int(indices.max()) >= I)):
                raise ValueError("%s: interval %d is corrupt" % (path, k))
            m = sp.csr_matrix((U, I), dtype=np.intc)
            # assign the arrays directly: no sort, no copy of the big arrays (they stay file-backed)
            # (scipy wants one index dtype: the row pointers -- U+1 entries -- are narrowed, indices are not copied)
            m.data, m.indices, m.indptr = data, indices, indptr.astype(np.int32)
            mats.append(m)
        off += sum(_align64(x) for x in sizes)
    return IntervalGraphs(U, I, mats, meta={"container": "SAGNNCSR v1", "mmap": bool(mmap)})


# --------------------------------------------------------------------------
# synthetic power-law interval graphs  
# --------------------------------------------------------------------------
# name -> (U, I, T, total_edges or per-interval list, L, d, alpha_user, alpha_item)
SHAPES = {
    "gowalla":     dict(U=48653, I=52621, T=3, E=1807125, L=2, d=64, au=0.8, ai=1.0),
    "amazon-book": dict(U=52643, I=91599, T=5, E=2984108, L=2, d=64, au=0.8, ai=1.0),
    "amazon-ref":  dict(U=11199, I=30821, T=5, E=[72280, 78997, 79692, 78096, 45651], L=3, d=64,
                        au=0.8, ai=1.0),
    "ml10m":       dict(U=69878, I=10677, T=6, E=10000054, L=3, d=128, au=0.6, ai=1.2),
    "scaled":      dict(U=10_000_000, I=2_000_000, T=8, E=1_000_000_000, L=2, d=64, au=0.8, ai=1.0),
    # the scaled config with U, I divided by 10 and E by 100 (same density): tables far beyond L2
    "scaled-s10":  dict(U=1_000_000, I=200_000, T=8, E=10_000_000, L=2, d=64, au=0.8, ai=1.0),
    # small cases for tests / smoke
    "tiny":        dict(U=300, I=200, T=3, E=3000, L=2, d=64, au=0.8, ai=1.0),
    "small":       dict(U=4000, I=3000, T=3, E=60000, L=2, d=64, au=0.8, ai=1.0),
}


def interval_sizes(total, T):
    """Equal shares, the last interval ~0.6x (mirrors Amazon's 45,651 vs ~78 K)."""
    w = np.ones(T)
    if T > 1:
        w[-1] = 0.6
    sizes = np.floor(total * w / w.sum()).astype(np.int64)
    sizes[0] += total - sizes.sum()
    return [int(s) for s in sizes]


def _zipf_cdf(n, alpha):
    p = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    c = np.cumsum(p)
    return c / c[-1]


def make_interval_graphs(U, I, T, E, au=0.8, ai=1.0, seed=100, **_unused):
    """Bipartite power-law interval graphs.

    User activity ~ rank^-au, item popularity ~ rank^-ai, ids scattered by a fixed
    random permutation, pairs de-duplicated globally so each (u,i) lives in exactly
    one interval, intervals sized by ``interval_sizes`` (or the explicit list ``E``);
    every interval gets one edge on the last user row and one on the last item
    column so the reference's pad-100 hack (model.py:87) is in range; stored values
    are int32 timestamps.  Canonical CSR (sorted, no duplicates) like
    ``csr_matrix((v,(r,c)))`` at preprocess_to_trnmat.ipynb:379.
    """
    rng = np.random.default_rng(seed)
    sizes = [int(e) for e in E] if isinstance(E, (list, tuple)) else interval_sizes(int(E), T)
    assert len(sizes) == T
    forced_per = 2 if (U > T and I > T) else 0
    sizes_rand = [max(0, s - forced_per) for s in sizes]
    total = int(sum(sizes_rand))
    if total > U * I // 2:
        raise ValueError("requested edges exceed half of the dense matrix")
    ucdf, icdf = _zipf_cdf(U, au), _zipf_cdf(I, ai)
    uperm, iperm = rng.permutation(U), rng.permutation(I)
    forced = set()
    if forced_per:
        for k in range(T):
            forced.add((U - 1) * I + k)           # (last user, item k)
            forced.add(k * I + (I - 1))           # (user k, last item)
    forced_keys = np.fromiter(forced, dtype=np.int64) if forced else np.zeros(0, np.int64)
    keys = np.zeros(0, dtype=np.int64)
    need = total
    rounds = 0
    while keys.size < total:
        n_draw = int(max(1024, (need * 1.3) + 1024))
        u = uperm[np.minimum(np.searchsorted(ucdf, rng.random(n_draw)), U - 1)]
        i = iperm[np.minimum(np.searchsorted(icdf, rng.random(n_draw)), I - 1)]
        new = u.astype(np.int64) * I + i
        keys = np.unique(np.concatenate([keys, new]))
        if forced_keys.size:
            keys = keys[~np.isin(keys, forced_keys)]
        need = total - keys.size
        rounds += 1
        if rounds > 200:
            raise RuntimeError("generator did not converge (graph too dense for the power law)")
    keys = rng.permutation(keys)[:total]
    sub, off = [], 0
    for k in range(T):
        kk = keys[off:off + sizes_rand[k]]
        off += sizes_rand[k]
        if forced_per:
            kk = np.concatenate([kk, np.array([(U - 1) * I + k, k * I + (I - 1)], dtype=np.int64)])
        kk = np.sort(kk)
        rows = (kk // I).astype(np.int32)
        cols = (kk % I).astype(np.int32)
        vals = np.random.default_rng(seed + 1000 + k).integers(TS_LO, TS_HI, size=kk.size, dtype=np.int64)
        m = sp.csr_matrix((vals.astype(np.intc), (rows, cols)), shape=(U, I))
        m.sort_indices()
        sub.append(m)
    return IntervalGraphs(U, I, sub, meta=dict(U=U, I=I, T=T, au=au, ai=ai, seed=seed))


def make_interval_device(U, I, E, au=0.8, ai=1.0, seed=100, device=None, **_unused):
    """ONE interval graph of a big shape, generated on the GPU (the scaled config's intervals have
    125 M edges: a minute each in numpy, a second here).  Same law as `make_interval_graphs` (user
    activity ~ rank^-au, item popularity ~ rank^-ai, ids scattered by a fixed bijection, pairs
    de-duplicated, one edge on the last user row and the last item column), but it keeps every distinct
    pair of ~1.1 E draws instead of trimming to exactly E -- the realised count is returned and must be
    reported.  Returns ``(row int32 [E'], col int32 [E'])`` CUDA tensors, row-major sorted, ready for
    ``build_plan([(row, col)], U, I)``."""
    import math
    import torch
    dev = torch.device(device if device is not None else "cuda")
    g = torch.Generator(device=dev).manual_seed(int(seed))

    def ranks(n, alpha, count):
        cdf = torch.cumsum(torch.arange(1, n + 1, device=dev, dtype=torch.float64).pow_(-alpha), 0)
        cdf /= cdf[-1].clone()
        out = torch.empty(count, dtype=torch.int64, device=dev)
        step = 1 << 25                                   # bounded scratch: 32 M draws at a time
        for o in range(0, count, step):
            m = min(step, count - o)
            out[o:o + m] = torch.searchsorted(cdf, torch.rand(m, generator=g, device=dev, dtype=torch.float64)).clamp_(max=n - 1)
        return out

    def scatter(r, n, mult):                             # fixed bijection rank -> id
        while math.gcd(mult, n) != 1:
            mult += 1
        return (r * mult + 12345) % n

    n_draw = int(E * 1.12) + 1024
    u = scatter(ranks(U, au, n_draw), U, 2654435761 % U or 1)
    i = scatter(ranks(I, ai, n_draw), I, 40503 % I or 1)
    key = u * I + i
    del u, i
    key = torch.cat([key, torch.tensor([(U - 1) * I, I - 1], device=dev, dtype=torch.int64)])
    key = torch.unique(key)                              # sorted: row-major order, duplicates dropped
    row = (key // I).to(torch.int32)
    col = (key % I).to(torch.int32)
    return row, col


def make_named(name, seed=100, scale=1.0):
    """Graphs for a named BASELINE shape; ``scale`` < 1 shrinks U, I and E together."""
    s = dict(SHAPES[name])
    if scale != 1.0:
        s["U"] = max(8, int(s["U"] * scale))
        s["I"] = max(8, int(s["I"] * scale))
        s["E"] = [max(4, int(e * scale)) for e in s["E"]] if isinstance(s["E"], list) else max(16, int(s["E"] * scale))
    g = make_interval_graphs(seed=seed, **s)
    g.meta.update(name=name, L=s["L"], d=s["d"], scale=scale)
    return g


def xavier_embeddings(T, rows, d, seed, dtype=np.float32):
    """``defineParam`` xavier-uniform for a 3-D shape (Utils/NNLayers.py:47-50):
    U(-a, a) with a = sqrt(6 / (T*(rows + d)))  (SURVEY 8d)."""
    a = np.sqrt(6.0 / (T * (rows + d)))
    return np.random.default_rng(seed).uniform(-a, a, size=(T, rows, d)).astype(dtype)
