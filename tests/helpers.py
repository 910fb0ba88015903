"""Shared test helpers (graph builders, oracle adapters)."""
import numpy as np
import scipy.sparse as sp

from oracle import propagate_oracle as po


def random_interval_mats(T, U, I, nnz, seed=0, timestamps=True):
    rng = np.random.default_rng(seed)
    mats = []
    for _ in range(T):
        n = min(nnz, U * I)
        keys = rng.choice(U * I, size=n, replace=False)
        r, c = keys // I, keys % I
        v = rng.integers(1388534400, 1406073600, size=n) if timestamps else np.ones(n)
        mats.append(sp.csr_matrix((v.astype(np.intc), (r, c)), shape=(U, I)))
    return mats


def adj_lists(mats):
    adj = [po.trans_to_lsts(m)[0] for m in mats]
    tp = [po.trans_to_lsts(po.transpose(m))[0] for m in mats]
    return adj, tp


def random_tables(T, U, I, d, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    return tuple((rng.standard_normal((T, n, d)) * scale).astype(np.float32) for n in (U, I, U, I))


def decode_gpu_masks(masks, T, U, I, d, L):
    """uint8 device buffer written by sagnn_propagate_fwd -> uint8 [T, L, (U+I)*d] in the C
    oracle's layout (1 = gradient passes unscaled).  Device layout per layer: bytes
    [T,U,d/4] then [T,I,d/4]; bit i (i < 4) of byte j of a row is element 4*j + i."""
    raw = masks.detach().cpu().numpy().view(np.uint8)
    bpr = d // 4
    per_layer = T * (U + I) * bpr
    out = np.empty((T, L, (U + I) * d), dtype=np.uint8)
    for l in range(L):
        lay = raw[l * per_layer:(l + 1) * per_layer]
        bits = np.unpackbits(lay.reshape(-1, 1), axis=1, bitorder="little")[:, :4].reshape(-1)
        out[:, l, :U * d] = bits[:T * U * d].reshape(T, U * d)
        out[:, l, U * d:] = bits[T * U * d:].reshape(T, I * d)
    return out


class DenseRowBackend:
    """CPU stand-in (tests only) for the C-ABI calls a row-sharded step makes
    (sagnn_propagate_fwd_layers / _bwd_levels / sagnn_workspace_table): same buffers, same wiring
    (ping-pong layer tables, pre-masked backward sources, owned rows only written), dense numpy
    arithmetic in the tensors' dtype.  Lets the gloo tests drive sagnn_b200.dist's real schedule."""

    def __init__(self, rs, d, mats):
        import torch
        self.torch = torch
        self.L, self.leaky, self.d = rs.n_layers, rs.leaky, d
        self.ub, self.ue, self.ib, self.ie = rs.row_block
        self.T = len(mats)
        self.A = []
        for m in mats:
            a = np.zeros((rs.U_pad, rs.I_pad))
            a[:m.shape[0], :m.shape[1]] = (m.toarray() != 0)
            self.A.append(a)
        z = lambda rows: torch.full((self.T, rows, d), float("nan"), dtype=torch.float64)
        self.buf = [(z(rs.U_pad), z(rs.I_pad)) for _ in range(2)]
        self.pm = [(z(rs.U_pad), z(rs.I_pad)) for _ in range(2)]
        self.masks = None

    def begin_forward(self, u, i, u_out, i_out, need_bwd):
        self.u, self.i, self.u_out, self.i_out = u, i, u_out, i_out
        self.masks = [(np.zeros(tuple(u.shape), bool), np.zeros(tuple(i.shape), bool)) for _ in range(self.L)]

    def begin_backward(self, gu, gi, d_u, d_i):
        self.gu, self.gi, self.d_u, self.d_i = gu, gi, d_u, d_i

    def _own(self, side):
        return slice(self.ub, self.ue) if side == 0 else slice(self.ib, self.ie)

    def _spmm(self, k, side, src):                      # owned rows of A_k (side 0) / A_k^T (side 1) times src[k]
        a = self.A[k] if side == 0 else self.A[k].T
        return a[self._own(side)] @ src[k].numpy()

    def fwd_layers(self, a, b):
        tn = self.torch.from_numpy
        for l in range(a, b):
            cur = (self.u, self.i) if l == 0 else self.buf[(l - 1) & 1]
            last = l == self.L - 1
            for side in (0, 1):
                own = self._own(side)
                out = (self.u_out, self.i_out)[side]
                for k in range(self.T):
                    z = self._spmm(k, side, cur[side ^ 1])
                    e = cur[side][k, own].numpy()
                    nxt = e + np.maximum(self.leaky * z, z)
                    self.masks[l][side][k, own] = ~(self.leaky * z >= z)
                    if not last:
                        self.buf[l & 1][side][k, own] = tn(nxt)
                    if last or l >= 1:
                        prev = None if l == 0 else ((self.u, self.i)[side] if l == 1 else out)
                        o = e if prev is None else prev[k, own].numpy() + e
                        out[k, own] = tn(o + nxt if last else o)

    def _src_index(self, step):
        return (1 if self.L >= 2 else 0) if step == 0 else (step - 1) & 1

    def bwd_levels(self, a, b):
        tn = self.torch.from_numpy
        G = (self.gu, self.gi)
        for ph in range(a, b):
            if ph == 0:
                for side in (0, 1):
                    own = self._own(side)
                    m = self.masks[self.L - 1][side][:, own]
                    g = G[side][:, own].numpy()
                    self.pm[self._src_index(0)][side][:, own] = tn(np.where(m, g, self.leaky * g))
                continue
            step, l = ph - 1, self.L - ph
            src = self.pm[self._src_index(step)]
            g = G if step == 0 else self.buf[(step - 1) & 1]
            o1 = (self.d_u, self.d_i) if l == 0 else self.buf[step & 1]
            for side in (0, 1):
                own = self._own(side)
                for k in range(self.T):
                    n = G[side][k, own].numpy() + g[side][k, own].numpy() + self._spmm(k, side, src[side ^ 1])
                    o1[side][k, own] = tn(n)
                    if l > 0:
                        m = self.masks[l - 1][side][k, own]
                        self.pm[step & 1][side][k, own] = tn(np.where(m, n, self.leaky * n))

    def table(self, which, index):
        return self.buf[index & 1] if which == 0 else self.pm[self._src_index(index)]
