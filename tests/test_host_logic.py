"""CPU tests: host-side mirror of the reference's data layer, the generator, the C-ABI surface."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import sagnn_b200 as sg
from sagnn_b200 import data_handler as dh, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_mat(z):
    return sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "index_*.npz"))))
def test_host_mirror_matches_reference_fixtures(path):
    """sagnn_b200.transToLsts / transpose == the reference's DataHandler functions (golden)."""
    z = np.load(path)
    m = _load_mat(z)
    for norm, tag in ((False, "raw"), (True, "norm")):
        idx, dat, shp = sg.transToLsts(m, norm=norm)
        np.testing.assert_array_equal(idx, z[f"adj_idx_{tag}"])
        np.testing.assert_array_equal(dat, z[f"adj_data_{tag}"])
        assert idx.dtype == np.int32 and dat.dtype == np.int32 and shp == list(z[f"adj_shape_{tag}"])
        tidx, tdat, _ = sg.transToLsts(sg.transpose(m), norm=norm)
        np.testing.assert_array_equal(tidx, z[f"tp_idx_{tag}"])
        np.testing.assert_array_equal(tdat, z[f"tp_data_{tag}"])


def test_trn_mat_time_round_trip(tmp_path):
    g = dh.make_named("tiny", seed=3)
    p = tmp_path / "trn_mat_time"
    dh.write_trn_mat_time(str(p), g)
    h = dh.load_trn_mat_time(str(p))
    assert (h.n_user, h.n_item, h.graph_num) == (g.n_user, g.n_item, 3)
    for a, b in zip(g.sub_mat, h.sub_mat):
        assert b.dtype == np.intc and (a != b).nnz == 0
    # trnMat[0] = interaction counts, timeMat = last interval id (interval 0 vanishes)
    assert h.trn_mat.shape == (g.n_user, g.n_item) and h.trn_mat.nnz == sum(g.nnz)
    assert h.time_mat.nnz == sum(g.nnz[1:])
    assert dh.load_trn_mat_time(str(p), graph_num=2).graph_num == 2
    with pytest.raises(IndexError):
        dh.load_trn_mat_time(str(p), graph_num=4)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "trnmat_*.npz"))))
def test_trans_sub_matches_reference_notebook_fixtures(path):
    """The producer of trn_mat_time: the vectorised `trans` / `trans_sub` mirror vs the
    outputs of the reference notebook's own cells (tests/golden/make_golden_trnmat.py), bit-exact:
    global count matrix, min / max timestamps, every interval graph (structure AND first-occurrence
    timestamps), timeMat with its dropped zeros."""
    z = np.load(path)
    U, I, T = int(z["U"]), int(z["I"]), int(z["T"])
    trn, minn, maxx = dh.trans(z["u"], z["i"], z["t"], U, I)
    assert (minn, maxx) == (int(z["minn"]), int(z["maxx"]))
    trn.sum_duplicates(); trn.sort_indices()
    assert np.array_equal(trn.indptr, z["trn_indptr"]) and np.array_equal(trn.indices, z["trn_indices"])
    assert np.array_equal(trn.data, z["trn_data"])
    subs, tm = dh.trans_sub(z["u"], z["i"], z["t"], U, I, T, minn, maxx)
    assert len(subs) == T
    for k, m in enumerate(subs):
        assert m.dtype == np.intc and m.has_canonical_format
        assert np.array_equal(m.indptr, z["sub%d_indptr" % k]) and np.array_equal(m.indices, z["sub%d_indices" % k])
        assert np.array_equal(m.data, z["sub%d_data" % k])
    tm.sort_indices()
    assert np.array_equal(tm.indptr, z["tm_indptr"]) and np.array_equal(tm.indices, z["tm_indices"])
    assert np.array_equal(tm.data, z["tm_data"])
    g = dh.make_trn_mat_time(z["u"], z["i"], z["t"], U, I, T)
    assert g.graph_num == T and g.nnz == [int(z["sub%d_indptr" % k][-1]) for k in range(T)]
    # each (user, item) pair can sit in several intervals, but at most once per interval
    assert all(m.nnz == len(set(zip(*m.nonzero()))) for m in g.sub_mat)


def test_interaction_to_triples_order():
    inter = [None, {5: [30, 10], 2: [20]}, {}, {1: None, 7: [40]}]
    u, i, t = dh.interaction_to_triples(inter)
    assert u.tolist() == [1, 1, 1, 3] and i.tolist() == [5, 5, 2, 7] and t.tolist() == [30, 10, 20, 40]
    with pytest.raises(ZeroDivisionError):
        dh.trans_sub([0], [0], [5], 2, 2, 3, 5, 5)


def test_binary_csr_container_round_trip(tmp_path):
    """the mmap-able container holds exactly what trn_mat_time[1] holds -- same canonical
    CSR, same intc values, same transToLsts / transpose outputs -- including an empty interval, the
    --graphNum prefix, the value-less variant and rejection of foreign / truncated files."""
    g = dh.make_named("tiny", seed=9)
    g.sub_mat[1] = sp.csr_matrix(g.sub_mat[1].shape, dtype=np.intc)            # an empty interval
    p = str(tmp_path / "trn_mat_time.sagnncsr")
    dh.write_trn_mat_bin(p, g)
    for mm in (True, False):
        h = dh.load_trn_mat_bin(p, mmap=mm)
        assert (h.n_user, h.n_item, h.graph_num, h.nnz) == (g.n_user, g.n_item, g.graph_num, g.nnz)
        for a, b in zip(g.sub_mat, h.sub_mat):
            assert b.dtype == np.intc and b.has_canonical_format
            assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            assert np.array_equal(a.data, b.data)
            for x, y in zip(dh.transToLsts(a), dh.transToLsts(b)):
                assert np.array_equal(np.asarray(x), np.asarray(y))
            assert (dh.transpose(a) != dh.transpose(b)).nnz == 0
    assert isinstance(dh.load_trn_mat_bin(p).sub_mat[0].indices, np.memmap)     # file-backed, not copied
    assert dh.load_trn_mat_bin(p, graph_num=2).graph_num == 2
    with pytest.raises(IndexError):
        dh.load_trn_mat_bin(p, graph_num=4)
    # same content as the pickle path
    pk = str(tmp_path / "trn_mat_time")
    dh.write_trn_mat_time(pk, g)
    for a, b in zip(dh.load_trn_mat_time(pk).sub_mat, dh.load_trn_mat_bin(p).sub_mat):
        assert (a != b).nnz == 0
    # structure only
    dh.write_trn_mat_bin(p, g.sub_mat, with_values=False)
    h = dh.load_trn_mat_bin(p)
    assert np.array_equal(h.sub_mat[0].indices, g.sub_mat[0].indices) and (h.sub_mat[0].data == 1).all()
    # foreign and truncated files
    bad = tmp_path / "bad"
    bad.write_bytes(b"not a container" * 8)
    with pytest.raises(ValueError):
        dh.load_trn_mat_bin(str(bad))
    whole = open(p, "rb").read()
    bad.write_bytes(whole[:len(whole) // 2])
    with pytest.raises(ValueError):
        dh.load_trn_mat_bin(str(bad))


def test_generator_properties():
    g = dh.make_named("small", seed=100)
    U, I = g.n_user, g.n_item
    assert g.nnz == dh.interval_sizes(60000, 3) and g.nnz[-1] < g.nnz[0]
    seen = sp.csr_matrix((U, I), dtype=np.int32)
    for m in g.sub_mat:
        assert m.has_canonical_format and m.dtype == np.intc
        assert m.data.min() >= dh.TS_LO and m.data.max() < dh.TS_HI
        idx, _, _ = sg.transToLsts(m)
        assert idx[-1, 0] + 101 >= U                     # pad-100 hack in range (model.py:87)
        assert sg.transToLsts(sg.transpose(m))[0][-1, 0] + 101 >= I
        seen = seen + (m != 0).astype(np.int32)
    assert seen.max() == 1                               # every (u,i) lives in exactly one interval
    deg = np.sort(np.asarray((seen != 0).sum(axis=0)).ravel())[::-1]
    assert deg[0] > 20 * max(1, np.median(deg))          # power-law head
    g2 = dh.make_named("small", seed=100)
    assert all((a != b).nnz == 0 for a, b in zip(g.sub_mat, g2.sub_mat))   # seeded


def test_amazon_ref_exact_interval_sizes():
    g = dh.make_named("amazon-ref", seed=100)
    assert (g.n_user, g.n_item) == (11199, 30821)
    assert g.nnz == [72280, 78997, 79692, 78096, 45651]   # preprocess_to_trnmat.ipynb:1911-1920


def test_xavier_limits():
    e = dh.xavier_embeddings(3, 48653, 64, 100)
    a = np.sqrt(6.0 / (3 * (48653 + 64)))
    assert e.dtype == np.float32 and np.abs(e).max() <= a and np.abs(e).max() > 0.99 * a


# ---------------------------------------------------------------- C ABI surface (no compute)
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "sagnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sagnn_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    syms = _header_symbols()
    assert len(syms) >= 16
    assert os.path.exists(_lib.lib_path()), "build the library first: python sa-gnn_b200/build.py"
    lib = ctypes.CDLL(_lib.lib_path())
    for s in syms:
        assert hasattr(lib, s), "libsagnn_b200.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms              # the ctypes binding covers the whole header


def test_library_loads_and_reports_version():
    lib = sg.load_library()
    assert b"sm_100a" in lib.sagnn_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sg.build_plan([sp.csr_matrix(np.eye(3, dtype=np.intc))])
    # the C ABI itself also fails loudly instead of computing on the host
    h = ctypes.c_void_p()
    nnz = (ctypes.c_int64 * 1)(3)
    rc = sg.load_library().sagnn_plan_create(1, 3, 3, nnz, ctypes.byref(h))
    assert rc == 3 and b"no CPU fallback" in sg.load_library().sagnn_last_error()


def test_argument_validation_needs_no_device():
    """Size limits and malformed arguments are rejected by the C ABI before any CUDA call (status +
    message, no exception across the boundary): the reference's maximum sizes (int32 row space), a
    latdim that is not a multiple of 4, an empty time range, too many peers."""
    lib = sg.load_library()
    err = lambda: lib.sagnn_last_error().decode()
    h = ctypes.c_void_p()
    nnz = (ctypes.c_int64 * 4)(1, 1, 1, 1)
    assert lib.sagnn_plan_create(4, 2 ** 29, 2 ** 29, nnz, ctypes.byref(h)) == 1 and "2^31" in err()   # T*(U+I) rows
    assert lib.sagnn_plan_create(0, 3, 3, nnz, ctypes.byref(h)) == 1
    assert lib.sagnn_pair_scores_fwd(None, 64, None, 64, None, None, 0, 66, 1, 0.5, None, None) == 1 and "multiple of 4" in err()
    assert lib.sagnn_pair_scores_fwd(None, 64, None, 64, None, None, 0, 64, 7, 0.5, None, None) == 1 and "activation" in err()
    assert lib.sagnn_pair_scores_fwd(None, 64, None, 64, None, None, 0, 64, 1, 0.5, None, None) == 0          # n = 0: nothing to do
    out = (ctypes.c_int64 * 3)()
    assert lib.sagnn_bucket_events(None, None, None, 5, 10, 10, 3, 100, 100, None, None, None, out, None) == 1 and "maxx" in err()
    assert lib.sagnn_bucket_events(None, None, None, 0, 10, 10, 3, 0, 100, None, None, None, out, None) == 0 and list(out) == [0, 0, 0]
    ptrs = (ctypes.c_void_p * 17)()
    assert lib.sagnn_propagate_fwd_scatter(None, None, None, None, None, 2, 64, 0.5, None, None, 0, 17, 0, ptrs, ptrs, None) == 1
    assert "world" in err()
    assert lib.sagnn_plan_set_row_block(None, 0, 1, 0, 1) == 1
    n_out = ctypes.c_int64()
    assert lib.sagnn_sample_ssl_batch(None, 0, None, 4, 2, 0, None, None, None, ctypes.byref(n_out), None) == 1


def test_product_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "sa-gnn_b200", "**", "*.py"), recursive=True) + \
            glob.glob(os.path.join(ROOT, "sa-gnn_b200", "csrc", "*")) + [os.path.join(ROOT, "sagnn_b200.py")]:
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), path
        assert "sagnn_oracle" not in src, path


def test_near_gpu_restores_affinity_and_degrades_without_nvml():
    """hostmem.near_gpu: the with-block may narrow the thread's CPU affinity to the GPU's NUMA node (first-touch
    placement of pinned pages); it must always put the old affinity back, and without a GPU / NVML it is a no-op."""
    import os
    from sagnn_b200 import hostmem
    before = os.sched_getaffinity(0)
    info = {}
    with hostmem.near_gpu(0, info) as got:
        assert got is info and "bound" in info
        inside = os.sched_getaffinity(0)
        assert inside <= before or not info["bound"]
    assert os.sched_getaffinity(0) == before
    with pytest.raises(RuntimeError):
        with hostmem.near_gpu(0):
            raise RuntimeError("body failed")
    assert os.sched_getaffinity(0) == before


def test_load_trn_mat_time_matches_reference_load_data():
    """a1 pinned to the reference: tests/golden/make_golden_loaddata.py CALLED the reference's own
    ``DataHandler().LoadData()`` (DataHandler.py:85-129) on tests/golden/dataset_tiny/ (the reference's on-disk layout)
    and recorded what it keeps; the product's loader must keep the same interval matrices, shape and dtype."""
    gold = os.path.join(os.path.dirname(__file__), "golden")
    fx = np.load(os.path.join(gold, "loaddata_tiny.npz"))
    h = dh.load_trn_mat_time(os.path.join(gold, "dataset_tiny", "trn_mat_time"))
    assert (h.n_user, h.n_item, h.graph_num) == (int(fx["user"]), int(fx["item"]), int(fx["T"]))
    for k, m in enumerate(h.sub_mat):
        assert str(m.dtype) == str(fx["sub%d_dtype" % k])
        np.testing.assert_array_equal(m.indptr, fx["sub%d_indptr" % k])
        np.testing.assert_array_equal(m.indices, fx["sub%d_indices" % k])
        np.testing.assert_array_equal(m.data, fx["sub%d_data" % k])
    np.testing.assert_array_equal(h.time_mat.indices, fx["time_indices"])
    np.testing.assert_array_equal(h.time_mat.data, fx["time_data"])
    # --graphNum selects a prefix of the stored intervals (model.py:230-231)
    assert dh.load_trn_mat_time(os.path.join(gold, "dataset_tiny", "trn_mat_time"), graph_num=2).graph_num == 2
    # the adjacency lists of the loaded matrices = what prepareModel builds from handler.subMat (model.py:227-238)
    for m in h.sub_mat:
        idx, data, shape = dh.transToLsts(m, norm=True)
        assert idx.shape == (m.nnz, 2) and shape == [h.n_user, h.n_item] or tuple(shape) == (h.n_user, h.n_item)


def test_bench_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference`` (the CPU arm the driver runs next to the GPU arm): one JSON line with the
    contract's keys, the same ``config`` builder as the GPU arm, no GPU needed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "edge_traversals/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    cfg = line["config"]
    assert cfg["edge_traversals_per_step"] == 4 * cfg["layers"] * cfg["edges"] and "workload" in cfg
    assert abs(line["value"] - cfg["edge_traversals_per_step"] / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
