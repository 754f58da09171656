"""ORACLE (test infrastructure, not product code): numpy evaluation of the
lowered NLP tables -- parameter tape V(p), g, J, f, grad f and the Lagrangian
Hessian slots.  It restates what CasADi's expanded SX functions compute for the
reference (optilayer.py:49-60: nlpsol builds f, g, J_g, H_L evaluators) on this
framework's table format.  Only tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke() may import it.  parity unpinned: no CasADi/IPOPT
binary exists in this image, see DESIGN.md.
"""
import numpy as np


def eval_tape(tb, p):
    V = np.zeros(tb.n_v)
    V[0] = 1.0
    V[1:1 + tb.n_par] = p
    base = 1 + tb.n_par
    for e in range(len(tb.tape_func)):
        lo, hi = tb.tape_ptr[e], tb.tape_ptr[e + 1]
        acc = 0.0
        for t in range(lo, hi):
            f = tb.tape_fac[t]
            acc += tb.tape_coef[t] * V[f[0]] * V[f[1]] * V[f[2]] * V[f[3]]
        fn = tb.tape_func[e]
        if fn == 1:
            acc = 1.0 / acc
        elif fn == 2:
            acc = 1.0 if acc >= 0.0 else 0.0
        elif fn == 3:
            acc = 1.0 if acc > 0.0 else 0.0
        elif fn == 4:
            acc = np.sin(acc)
        elif fn == 5:
            acc = np.cos(acc)
        elif fn == 6:
            acc = np.sqrt(acc)
        V[base + e] = acc
    return V


def _eval_terms(tl, V, xe, lam_ext=None):
    vals = tl.coef * V[tl.cidx]
    for k in range(tl.xi.shape[1]):
        vals = vals * xe[tl.xi[:, k]]
    if lam_ext is not None:
        vals = vals * lam_ext[tl.lrow]
    out = np.zeros(tl.n_out)
    seg = np.repeat(np.arange(tl.n_out), np.diff(tl.ptr))
    np.add.at(out, seg, vals)
    return out


class TableEval(object):
    """Intermediates ('mid' symbols, lowering.py): x_ext = [x, 1, mids]; the
    mids are the pseudo-rows m.. of G.  Rows are affine in the mids, so
    J = J_direct + A C (A = d row/d mid, C = d mid/d x, both slots of the J
    term list) and the Lagrangian Hessian takes mu = A^T lam as the
    multipliers of the pseudo-rows.  When A depends on x or on other mids the W
    term list has nnz_wx extra slots (X[l,k] = sum_i lam_i d2 row_i/d mid_l d x_k,
    M[l1,l2] likewise for two mids) and the Hessian gains X^T C + C^T X + C^T M C
    through the product lists xq_*."""

    def __init__(self, tb):
        self.tb = tb
        self.n_mid = getattr(tb, 'n_mid', 0)

    def tape(self, p):
        return eval_tape(self.tb, np.asarray(p, dtype=float))

    def _xe(self, x, V=None):
        xe = np.r_[np.asarray(x, dtype=float), 1.0, np.zeros(self.n_mid)]
        if self.n_mid and V is not None:
            tb = self.tb
            lo = tb.G.ptr[tb.m]
            sub = _Slice(tb.G, lo, len(tb.G.coef), tb.m, tb.m + self.n_mid)
            xe[tb.n + 1:] = _eval_terms(sub, V, xe)
        return xe

    def g(self, x, V):
        return _eval_terms(self.tb.G, V, self._xe(x, V))[:self.tb.m]

    def f(self, x, V):
        return _eval_terms(self.tb.F, V, self._xe(x))[0]

    def gradf(self, x, V):
        return _eval_terms(self.tb.DF, V, self._xe(x))

    def _jac_all(self, x, V):
        return _eval_terms(self.tb.J, V, self._xe(x, V))

    def jac_vals(self, x, V):
        tb = self.tb
        jv = self._jac_all(x, V)
        if not self.n_mid:
            return jv
        out = jv[:tb.nnz_j].copy()
        seg = np.repeat(np.arange(tb.nnz_j), np.diff(tb.jp_ptr))
        np.add.at(out, seg, jv[tb.jp_a] * jv[tb.jp_c])
        return out

    def jac_dense(self, x, V):
        J = np.zeros((self.tb.m, self.tb.n))
        J[self.tb.jrow, self.tb.jcol] = self.jac_vals(x, V)
        return J

    def hess_vals(self, x, V, lam, obj_factor=1.0):
        tb = self.tb
        lam = np.asarray(lam, dtype=float)
        lam_ext = np.r_[lam, obj_factor]
        if self.n_mid:
            jv = self._jac_all(x, V)
            mu = np.zeros(self.n_mid)
            seg = np.repeat(np.arange(self.n_mid), np.diff(tb.mu_ptr))
            np.add.at(mu, seg, lam[tb.mu_row] * jv[tb.mu_slot])
            lam_ext = np.r_[lam_ext, mu]
        return _eval_terms(tb.W, V, self._xe(x, V), lam_ext)

    def hess_cross(self, x, V, wvals):
        """(H position, value) of the contributions X^T C + C^T X + C^T M C."""
        tb = self.tb
        jv = self._jac_all(x, V)
        second = np.where(tb.xq_b >= 0, jv[np.maximum(tb.xq_b, 0)], 1.0)
        prod = wvals[tb.nnz_w + tb.xq_w] * jv[tb.xq_a] * second
        out = np.zeros(tb.n_xq)
        np.add.at(out, np.repeat(np.arange(tb.n_xq), np.diff(tb.xq_ptr)), prod)
        return tb.xq_h, out

    def hess_dense(self, x, V, lam, obj_factor=1.0):
        tb = self.tb
        W = np.zeros((tb.n, tb.n))
        vals = self.hess_vals(x, V, lam, obj_factor)
        W[tb.wrow, tb.wcol] = vals[:tb.nnz_w]
        if getattr(tb, 'nnz_wx', 0):
            q, v = self.hess_cross(x, V, vals)
            np.add.at(W, (tb.hrow[q], tb.hcol[q]), v)
        W = W + np.tril(W, -1).T
        return W


class _Slice(object):
    """Term-list view of the slots [s0, s1) (terms [lo, hi))."""

    def __init__(self, tl, lo, hi, s0, s1):
        self.coef, self.cidx = tl.coef[lo:hi], tl.cidx[lo:hi]
        self.xi, self.lrow = tl.xi[lo:hi], tl.lrow[lo:hi]
        self.ptr = tl.ptr[s0:s1 + 1] - lo
        self.n_out = s1 - s0
