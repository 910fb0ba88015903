"""Generates tests/golden/downstream_*.npz by EXECUTING the reference's own model.py text for the rows right
AFTER the propagation path.  Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden_downstream.py

Same mechanism as make_golden_model.py: ``tests/golden/tf1_shim.py`` stands in for TensorFlow 1.14, ``import model``
imports /root/reference/model.py UNMODIFIED (with the reference's own Params, Utils.NNLayers, Utils.attention), and
these line ranges of ``Recommender.ours()`` are exec'd as they stand, fed with a given ``user_vector`` /
``item_vector`` (the [T,R,d] stacks of model.py:131-132, i.e. the path's output):

  * model.py:133-156  the [R,T,d] transposes, ONE BasicLSTMCell (DropoutWrapper, MultiRNNCell) run by two
    ``tf.nn.dynamic_rnn`` calls, layer norm, ``MultiHeadSelfAttention.attention`` (Utils/attention.py:46-78 with
    ``ScaledDotProductAttention`` :31-44, the reference's own classes), mean over the interval axis
    -> ``final_user_vector`` / ``final_item_vector``                                             (N1)
  * model.py:111-112, 157-168  position embedding, the masked sums of the sequence's item vectors and positions,
    ``args.att_layer`` rounds of layer norm + ``MultiHeadSelfAttention`` + LeakyReLU residual -> ``att_user``
  * model.py:169-173  ``preds`` = dot product of the final vectors + the sequence-attention term               (N2)
    (``preds_dot`` in the fixtures is the first half alone, model.py:169-172)
  * model.py:241-244  (prepareModel) the prediction hinge ``preLoss`` over positives | negatives of ``preds``
  * model.py:174-203  the meta-weight network (``FC`` of Utils/NNLayers.py), ``preds_one`` per interval on
    ``user_vector[i]`` / ``item_vector[i]``, the weighted hinge ``sslloss``                        (N2)

The text runs twice: once to discover the variables it creates (TF initial values: glorot kernels, zero biases,
beta 0 / gamma 1), then -- after moving every variable off its trivial value -- in "replay" mode with the stored
values, in fp64 and in fp32.  What comes from the reference text: wiring, which tensors feed which op, the shared
LSTM cell, separate Q / K / V kernels per side, exp-normalised attention with 1e-8, axis choices, stop_gradient
placement, slicing of positives / negatives.  What the shim states: the per-op semantics of the TF kernels (gate
order and forget bias of BasicLSTMCell, layer_norm's axes and epsilon, dense = kernel + bias).  The input gradient
of the fusion is recorded as fp64 central differences of the executed text (rows are independent, so one
perturbation moves the same (interval, column) coordinate of every row).
"""
import os

import numpy as np

from make_golden_model import HERE, load_reference, ref_block


def blocks():
    return {
        "pos": ref_block(111, 112, ["posEmbed=NNs.defineParam('posEmbed', [args.pos_length, args.latdim], reg=True)",
                                    "pos= tf.tile(tf.expand_dims(tf.range(args.pos_length),axis=0),[args.batch,1])"]),
        "sequence": ref_block(157, 168, ["self.multihead_self_attention_sequence.append(MultiHeadSelfAttention(args.latdim,args.num_attention_heads))",
                                         "sequence_batch=tf.contrib.layers.layer_norm(tf.matmul(tf.expand_dims(self.mask,axis=1),tf.nn.embedding_lookup(iEmbed_att,self.sequence)))",
                                         "sequence_batch+=tf.contrib.layers.layer_norm(tf.matmul(tf.expand_dims(self.mask,axis=1),tf.nn.embedding_lookup(posEmbed,pos)))",
                                         "att_layer=Activate(att_layer1,\"leakyRelu\")+att_layer",
                                         "att_user=tf.reduce_sum(att_layer,axis=1)"]),
        "preloss": ref_block(241, 244, ["sampNum = tf.shape(self.uids)[0] // 2",
                                        "self.preLoss = tf.reduce_mean(tf.maximum(0.0, 1.0 - (self.posPred - self.negPred)))"]),
        "preds_full": ref_block(173, 173, ["preds += tf.reduce_sum(Activate(tf.nn.embedding_lookup(att_user,self.uLocs_seq),\"leakyRelu\")* pckIlat_att,axis=-1)"]),
        "fusion": ref_block(133, 156, ["user_vector_tensor=tf.transpose(user_vector, perm=[1, 0, 2])",
                                       "return tf.contrib.rnn.BasicLSTMCell(args.latdim)",
                                       "rnn_cell = tf.contrib.rnn.MultiRNNCell(cells, state_is_tuple=True)",
                                       "item_vector_rnn, _ = tf.nn.dynamic_rnn(cell=rnn_cell, inputs=item_vector_tensor",
                                       "self.multihead_self_attention0.attention(tf.contrib.layers.layer_norm(user_vector_tensor))",
                                       "final_item_vector = tf.reduce_mean(multihead_item_vector,axis=1)",
                                       "iEmbed_att=final_item_vector"]),
        "preds": ref_block(169, 172, ["pckUlat = tf.nn.embedding_lookup(final_user_vector, self.uids)",
                                      "preds = tf.reduce_sum(pckUlat * pckIlat, axis=-1)"]),
        "ssl": ref_block(174, 203, ["meta1=tf.concat([final_user_vector*user_vector[i],final_user_vector,user_vector[i]],axis=-1)",
                                    'activation=\'sigmoid\',reg=True,reuse=True,name="meta3"',
                                    "sampNum = tf.shape(self.suids[i])[0] // 2",
                                    "posPred_final = tf.stop_gradient(tf.slice(S_final, [0], [sampNum]))",
                                    "S_final = posweight_final*posPred_final-negweight_final*negPred_final",
                                    "preds_one = tf.reduce_sum(Activate(pckUlat* pckIlat , self.actFunc), axis=-1)",
                                    "sslloss += tf.reduce_sum(tf.maximum(0.0, 1.0 -S_final * (posPred-negPred)))",
                                    "self.preds_one.append(preds_one)"]),
    }


class _Rec:
    pass


ALL = ("pos", "fusion", "sequence", "preds", "preds_full", "ssl")


def run(shim, model, NNs, blk, uv, iv, ids, leaky, keep=1.0, which=ALL):
    """Executes the blocks on user_vector / item_vector [T,R,d]; returns the namespace and the Recommender stand-in."""
    NNs.params.clear(); NNs.regParams.clear()
    NNs.leaky = leaky
    rec = _Rec()
    rec.keepRate, rec.actFunc = keep, "leakyRelu"
    if ids is not None:
        rec.uids, rec.iids = shim.Tensor(ids["uids"]), shim.Tensor(ids["iids"])
        rec.suids = [shim.Tensor(a) for a in ids["suids"]]
        rec.siids = [shim.Tensor(a) for a in ids["siids"]]
        rec.sequence, rec.uLocs_seq = shim.Tensor(ids["sequence"]), shim.Tensor(ids["uLocs_seq"])
        rec.mask = shim.Tensor(ids["mask"].astype(uv.dtype))          # a float32 placeholder in the reference (fed with 0 / 1)
    ns = dict(model.__dict__)
    ns.update(self=rec, user_vector=shim.Tensor(uv), item_vector=shim.Tensor(iv))
    for name in which:
        exec(blk[name][0], ns)
        if name == "preds":
            ns["preds_dot"] = ns["preds"]                             # model.py:172, before :173 adds the sequence term
    return ns, rec


def main():
    shim, model, NNs = load_reference()
    args = model.args
    blk = blocks()
    rng = np.random.default_rng(20261019)
    cases = {
        # name: (T, U, I, d, heads, ssldim, leaky, pairs per interval, prediction pairs, att_layer, args.batch, pos_length)
        "a_t3_d64_h16_gowalla_sh": (3, 30, 24, 64, 16, 32, 0.5, 40, 50, 2, 12, 9),   # gowalla.sh: T=3, d=64, 16 heads, ssldim 32
        "b_t5_d32_h8": (5, 19, 26, 32, 8, 16, 0.1, 14, 21, 1, 7, 5),
    }
    for name, (T, U, I, d, heads, ssldim, leaky, npair, npred, att_layer, abatch, plen) in cases.items():
        args.user, args.item, args.latdim, args.graphNum = U, I, d, T
        args.num_attention_heads, args.ssldim, args.leaky = heads, ssldim, leaky
        args.att_layer, args.batch, args.pos_length = att_layer, abatch, plen
        uv = (0.5 * rng.standard_normal((T, U, d))).astype(np.float32)
        iv = (0.5 * rng.standard_normal((T, I, d))).astype(np.float32)
        ids = dict(uids=rng.integers(0, U, size=2 * npred).astype(np.int32), iids=rng.integers(0, I, size=2 * npred).astype(np.int32),
                   suids=[np.tile(rng.integers(0, U, size=npair + k), 2).astype(np.int32) for k in range(T)],   # positives | negatives share the users (model.py:323-330 interleaves; trainEpoch feeds what the sampler returns)
                   siids=[rng.integers(0, I, size=2 * (npair + k)).astype(np.int32) for k in range(T)])
        # what sampleTrainBatch feeds (model.py:283-297): right-aligned item sequences, their 0 / 1 mask, and for every
        # prediction pair the batch position of its user
        seq = np.zeros((abatch, plen), np.int64); msk = np.zeros((abatch, plen), np.float32)
        for b in range(abatch - 2):                                   # the last two rows stay padding (batch < args.batch)
            n = int(rng.integers(1, plen + 1))
            seq[b, plen - n:] = rng.integers(0, I, size=n); msk[b, plen - n:] = 1
        ids.update(sequence=seq, mask=msk, uLocs_seq=rng.integers(0, abatch - 2, size=2 * npred).astype(np.int32))
        # pass 1: discover the variables (TF initial values), then move them off 0 / 1 and round to fp32
        shim.VarStore.reset()
        run(shim, model, NNs, blk, uv, iv, ids, leaky)
        names = list(shim.VarStore.names)
        shim.VarStore.values = [(v + 0.1 * rng.standard_normal(np.shape(v))).astype(np.float32) for v in shim.VarStore.values]
        params = [v.copy() for v in shim.VarStore.values]
        # pass 2: replay in fp64 (the fixture) and fp32 (what the reference's dtype gives)
        outs = {}
        for tag, dt in (("", np.float64), ("_f32", np.float32)):
            shim.VarStore.values = [p.copy() for p in params]
            shim.VarStore.replay(dt)
            ns, rec = run(shim, model, NNs, blk, uv.astype(dt), iv.astype(dt), ids, leaky)
            assert shim.VarStore.cursor == len(names), "replay consumed a different number of variables"
            outs.update({"preds_dot" + tag: ns["preds_dot"].a, "att_user" + tag: ns["att_user"].a})
            rec.preds = ns["preds"]                                   # model.py:240: self.preds, self.sslloss = self.ours()
            exec(blk["preloss"][0], ns)                               # model.py:241-244 (prepareModel)
            outs["preLoss" + tag] = np.asarray(rec.preLoss.a)
            assert ns["final_user_vector"].a.dtype == dt
            outs.update({"final_user_vector" + tag: ns["final_user_vector"].a, "final_item_vector" + tag: ns["final_item_vector"].a,
                         "preds" + tag: ns["preds"].a, "user_weight" + tag: ns["user_weight"].a,
                         "sslloss" + tag: np.asarray(ns["sslloss"].a)})
            for k in range(T):
                outs["preds_one%d%s" % (k, tag)] = rec.preds_one[k].a
        # the same text with the dropout wrapper active must differ (the wrapper really sits on the LSTM output)
        shim.VarStore.values = [p.copy() for p in params]; shim.VarStore.replay(np.float64)
        nsd, _ = run(shim, model, NNs, blk, uv.astype(np.float64), iv.astype(np.float64), None, leaky, keep=0.5, which=("pos", "fusion"))
        assert not np.allclose(nsd["final_user_vector"].a, outs["final_user_vector"])
        # input gradient of the fusion: d/d(user_vector, item_vector) of sum(wu * final_user) + sum(wi * final_item)
        wu = rng.standard_normal((U, d)); wi = rng.standard_normal((I, d))

        def loss_rows(u, i):
            shim.VarStore.values = [p.copy() for p in params]; shim.VarStore.replay(np.float64)
            n, _ = run(shim, model, NNs, blk, u, i, None, leaky, which=("pos", "fusion"))
            return (wu * n["final_user_vector"].a).sum(axis=1), (wi * n["final_item_vector"].a).sum(axis=1)

        h = 1e-5
        u0, i0 = uv.astype(np.float64), iv.astype(np.float64)
        d_uv, d_iv = np.zeros_like(u0), np.zeros_like(i0)
        for t in range(T):
            for c in range(d):
                up, ip = u0.copy(), i0.copy(); up[t, :, c] += h; ip[t, :, c] += h
                um, im = u0.copy(), i0.copy(); um[t, :, c] -= h; im[t, :, c] -= h
                (lup, lip), (lum, lim) = loss_rows(up, ip), loss_rows(um, im)
                d_uv[t, :, c] = (lup - lum) / (2 * h)
                d_iv[t, :, c] = (lip - lim) / (2 * h)
        out = dict(T=T, U=U, I=I, d=d, heads=heads, ssldim=ssldim, leaky=leaky, user_vector=uv, item_vector=iv,
                   att_layer=att_layer, batch=abatch, pos_length=plen, sequence=ids["sequence"], mask=ids["mask"],
                   uLocs_seq=ids["uLocs_seq"],
                   uids=ids["uids"], iids=ids["iids"], w_user=wu, w_item=wi, d_user_vector=d_uv, d_item_vector=d_iv,
                   var_names=np.array(names), **outs)
        for k in range(T):
            out["suids%d" % k], out["siids%d" % k] = ids["suids"][k], ids["siids"][k]
        for j, p in enumerate(params):
            out["var%02d" % j] = p
        np.savez_compressed(os.path.join(HERE, "downstream_%s.npz" % name), **out)
        print("downstream", name, "variables:", [n.split("#")[0] for n in names])
        print("   final_user |max|", np.abs(outs["final_user_vector"]).max(), "sslloss", float(outs["sslloss"]),
              "fp32 vs fp64 final_user", np.abs(outs["final_user_vector_f32"] - outs["final_user_vector"]).max())


if __name__ == "__main__":
    main()
