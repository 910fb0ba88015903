"""Generates tests/golden/modelref_*.npz by EXECUTING the reference's own model.py text.
Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_model.py

TensorFlow 1.14 cannot be installed here, so ``tests/golden/tf1_shim.py`` is registered under the
name ``tensorflow`` (numpy-backed eager stand-ins for the ~15 TF ops the path uses) and then

  * ``import model`` imports /root/reference/model.py UNMODIFIED (with the reference's own
    Params, Utils.NNLayers, DataHandler);
  * the adjacency lists are built by the reference's ``prepareModel`` lines 227-238, exec'd as they
    stand (``transToLsts`` / ``transpose`` are the reference's functions);
  * the propagation is the reference's ``ours()`` lines 105, 113-114, 117-134, exec'd as they stand,
    which call the reference's ``Recommender.messagePropagate`` / ``edgeDropout`` methods
    (model.py:80-102) and ``Activate`` (Utils/NNLayers.py:150-...).

So wiring, source / target roles, the Jacobi update order, the pad-100 + identity lookup, the
LeakyReLU form, the layer sum and the stack / transpose hand-off come from the reference text;
only the per-op semantics of the TF kernels come from the shim.  Gradients are central finite
differences (fp64) of that executed forward -- exact for this piecewise-linear map as long as no
activation argument lies within the step of its kink, which the generator checks.

Outputs per case: inputs (adjacency lists, fp32 embeddings, fp32 upstream), ``user_vector`` /
``item_vector`` ([T,R,d], model.py:131-132) and ``user_vector_tensor`` / ``item_vector_tensor``
([R,T,d], model.py:133-134) in fp64 and fp32, ``dU`` / ``dI`` (fp64 finite differences).
``modelref_errors.npz`` records the inputs on which the reference graph cannot be built / run
(one-edge interval -> rank-0 segment ids; populated rows ending more than 100 before R).
"""
import hashlib
import os
import sys
import textwrap

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference():
    sys.path.insert(0, HERE)
    import tf1_shim
    tf1_shim.install()
    sys.argv = ["x"]                       # Params.py parses argv at import
    sys.path.insert(0, REF)
    import model                           # /root/reference/model.py, unmodified
    import Utils.NNLayers as NNs
    assert os.path.realpath(model.__file__) == os.path.join(REF, "model.py")
    return tf1_shim, model, NNs


def ref_block(first, last, must_contain):
    """Lines first..last of the reference's model.py as they stand (dedented for exec)."""
    with open(os.path.join(REF, "model.py")) as f:
        lines = f.read().split("\n")
    text = "\n".join(lines[first - 1:last])
    for s in must_contain:
        assert s in text, f"model.py:{first}-{last} no longer contains {s!r}: the reference moved"
    return compile(textwrap.dedent(text), f"/root/reference/model.py:{first}-{last}", "exec"), text


class _Handler:
    pass


def build(shim, model, NNs, mats, U, I, d, L, leaky, keep_rate):
    """Recommender with subAdj / subTpAdj built by model.py:227-238 from scipy matrices."""
    args = model.args
    args.user, args.item, args.latdim, args.graphNum, args.gnn_layer, args.leaky = U, I, d, len(mats), L, leaky
    NNs.leaky = args.leaky                 # model.py:208
    NNs.params.clear(); NNs.regParams.clear()
    rec = model.Recommender.__new__(model.Recommender)
    rec.actFunc = "leakyRelu"              # model.py:209
    rec.keepRate = keep_rate               # model.py:206 (a placeholder fed with args.keepRate / 1.0)
    rec.handler = _Handler()
    rec.handler.subMat = mats
    rec.handler.maxTime = 3
    code, text = ref_block(227, 238, ["transToLsts(seqadj, norm=True)", "transToLsts(transpose(seqadj), norm=True)",
                                       "self.subTpAdj.append", "self.maxTime=self.handler.maxTime"])
    ns = dict(model.__dict__); ns["self"] = rec
    exec(code, ns)
    return rec, text


_BLOCKS = None


def ours_blocks():
    global _BLOCKS
    if _BLOCKS is None:
        _BLOCKS = [ref_block(105, 105, ["user_vector,item_vector=list(),list()"]),
                   ref_block(113, 114, ["self.items=tf.range(args.item)", "self.users=tf.range(args.user)"]),
                   ref_block(117, 134, ["for k in range(args.graphNum):", "embs0=[uEmbed[k]]",
                                        "self.messagePropagate(embs1[-1],self.edgeDropout(self.subAdj[k]),'user')",
                                        "self.messagePropagate(embs0[-1],self.edgeDropout(self.subTpAdj[k]),'item')",
                                        "embs0.append(a_emb0+embs0[-1])", "user=tf.add_n(embs0)",
                                        "user_vector=tf.stack(user_vector,axis=0)",
                                        "item_vector_tensor=tf.transpose(item_vector, perm=[1, 0, 2])"])]
    return _BLOCKS


def run_ours(shim, model, NNs, rec, uE, iE):
    """model.py:105,113-114,117-134 with uEmbed / iEmbed given (the parameters of :108-109)."""
    NNs.params.clear(); NNs.regParams.clear()
    ns = dict(model.__dict__)
    ns.update(self=rec, uEmbed=shim.Tensor(uE), iEmbed=shim.Tensor(iE))
    for code, _ in ours_blocks():
        exec(code, ns)
    return ns["user_vector"].a, ns["item_vector"].a, ns["user_vector_tensor"].a, ns["item_vector_tensor"].a


def fd_grads(shim, model, NNs, mats, U, I, d, L, leaky, uE, iE, gU, gI, h=1e-8):
    """dU, dI = gradient of sum(gU*user_vector) + sum(gI*item_vector), central differences in fp64,
    one interval at a time (intervals are independent: model.py:118-129)."""
    dU = np.zeros(uE.shape, np.float64); dI = np.zeros(iE.shape, np.float64)
    for k in range(len(mats)):
        rec, _ = build(shim, model, NNs, [mats[k]], U, I, d, L, leaky, 1.0)
        base_u = uE[k:k + 1].astype(np.float64); base_i = iE[k:k + 1].astype(np.float64)
        g0 = gU[k].astype(np.float64); g1 = gI[k].astype(np.float64)

        def loss(u, i):
            uv, iv, _, _ = run_ours(shim, model, NNs, rec, u, i)
            return float(np.sum(g0 * uv[0]) + np.sum(g1 * iv[0]))

        for tab, out in ((base_u, dU), (base_i, dI)):
            flat = tab.reshape(-1)
            res = out[k].reshape(-1)
            for j in range(flat.size):
                old = flat[j]
                flat[j] = old + h; lp = loss(base_u, base_i)
                flat[j] = old - h; lm = loss(base_u, base_i)
                flat[j] = old
                res[j] = (lp - lm) / (2 * h)
    return dU, dI


def rand_mat(rng, U, I, dens, last_row=None, ts=True):
    m = sp.random(U, I, density=dens, random_state=int(rng.integers(1 << 30)), format="csr")
    m.data[:] = rng.integers(1388534400, 1406073600, size=m.nnz) if ts else 1
    m = m.astype(np.intc)
    if last_row is not None:               # rows >= last_row empty: exercises the pad-100 + lookup trick
        keep = sp.diags((np.arange(U) < last_row).astype(np.intc))
        m = sp.csr_matrix(keep @ m).astype(np.intc)
        m.eliminate_zeros()
    return m


def main():
    shim, model, NNs = load_reference()
    rng = np.random.default_rng(20261018)
    texts = []
    cases = {
        # name: (T, U, I, d, L, leaky, density, keepRate, empty tail rows in interval 1)
        "a_t2_l2_d32": (2, 40, 30, 32, 2, 0.5, 0.08, 1.0, None),
        "b_t3_l2_d64_gowalla_sh": (3, 60, 45, 64, 2, 0.5, 0.08, 0.5, 41),     # gowalla.sh: T=3, L=2, d=64, leaky 0.5
        "c_t1_l3_d128_leaky01": (1, 33, 21, 128, 3, 0.1, 0.15, 1.0, None),
    }
    for name, (T, U, I, d, L, leaky, dens, keep, tail) in cases.items():
        for attempt in range(50):          # redraw until no activation argument sits next to its kink
            mats = [rand_mat(rng, U, I, dens, last_row=(tail if k == 1 else None)) for k in range(T)]
            uE = rng.uniform(-1, 1, size=(T, U, d)).astype(np.float32)
            iE = rng.uniform(-1, 1, size=(T, I, d)).astype(np.float32)
            gU = rng.normal(size=(T, U, d)).astype(np.float32)
            gI = rng.normal(size=(T, I, d)).astype(np.float32)
            rec, t227 = build(shim, model, NNs, mats, U, I, d, L, leaky, keep)
            shim.Stats.reset()
            uv64, iv64, uvt64, ivt64 = run_ours(shim, model, NNs, rec, uE.astype(np.float64), iE.astype(np.float64))
            gap = shim.Stats.min_gap       # min |leaky*z - z| over all non-empty rows
            if gap > 2e-6:      # a step of 1e-8 moves an argument by < 1e-6
                break
        assert gap > 2e-6, f"{name}: an activation argument is within {gap} of its kink"
        uv32, iv32, uvt32, ivt32 = run_ours(shim, model, NNs, rec, uE, iE)
        assert uv32.dtype == np.float32
        if keep != 1.0:                    # edge dropout rewrites only the ignored values (model.py:93-102)
            rec1, _ = build(shim, model, NNs, mats, U, I, d, L, leaky, 1.0)
            uvk, ivk, _, _ = run_ours(shim, model, NNs, rec1, uE.astype(np.float64), iE.astype(np.float64))
            assert np.array_equal(uvk, uv64) and np.array_equal(ivk, iv64)
        dU, dI = fd_grads(shim, model, NNs, mats, U, I, d, L, leaky, uE, iE, gU, gI)
        out = dict(T=T, U=U, I=I, d=d, L=L, leaky=leaky, keepRate=keep, min_kink_gap=gap, uE=uE, iE=iE, gU=gU, gI=gI,
                   user_vector=uv64, item_vector=iv64, user_vector_tensor=uvt64, item_vector_tensor=ivt64,
                   user_vector_f32=uv32, item_vector_f32=iv32, dU=dU, dI=dI)
        for k in range(T):
            out[f"adj{k}"] = np.asarray(rec.subAdj[k].indices.a, np.int32)
            out[f"tp{k}"] = np.asarray(rec.subTpAdj[k].indices.a, np.int32)
            out[f"adj_values{k}"] = np.asarray(rec.subAdj[k].values.a, np.int32)
            out[f"csr_indptr{k}"] = mats[k].indptr
            out[f"csr_indices{k}"] = mats[k].indices
            out[f"csr_data{k}"] = mats[k].data
        np.savez_compressed(os.path.join(HERE, f"modelref_{name}.npz"), **out)
        print("modelref", name, "kink gap", gap, "max|dU|", np.abs(dU).max())
        texts.append(t227)

    # inputs on which the reference itself fails (recorded so the tests can state the deviation)
    errs = {}
    U, I, d = 30, 20, 32
    one = sp.csr_matrix((np.array([1400000000], np.intc), ([3], [2])), shape=(U, I))
    empty = sp.csr_matrix((U, I), dtype=np.intc)            # transToLsts substitutes the edge (0,0): also one edge
    far = rand_mat(rng, 250, 40, 0.05, last_row=100)         # last populated row + 101 < R
    for nm, m in (("one_edge", one), ("empty", empty), ("tail_gap_gt_100", far)):
        Uc, Ic = m.shape
        rec, _ = build(shim, model, NNs, [m], Uc, Ic, d, 1, 0.5, 1.0)
        try:
            run_ours(shim, model, NNs, rec, np.ones((1, Uc, d), np.float32), np.ones((1, Ic, d), np.float32))
            errs[nm] = "ran"
        except Exception as e:             # noqa: BLE001
            errs[nm] = f"{type(e).__name__}: {e}"
        print("reference on", nm, "->", errs[nm])
    src = "\n".join(t for _, t in ours_blocks()) + "\n".join(texts[:1])
    np.savez_compressed(os.path.join(HERE, "modelref_errors.npz"),
                        **{k: np.array(v) for k, v in errs.items()},
                        model_py_sha256=np.array(hashlib.sha256(open(os.path.join(REF, "model.py"), "rb").read()).hexdigest()),
                        executed_text=np.array(src))


if __name__ == "__main__":
    main()
