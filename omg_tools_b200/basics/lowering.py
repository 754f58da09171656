"""Lower a polynomial NLP (rows of ``Poly``) to the flat constant tables that
the CUDA kernels (csrc/omg_b200.cu) and the CPU oracle (oracle/) consume.

This replaces what the reference obtains from CasADi: ``nlpsol(..., {x,p,f,g},
{expand: True})`` builds SX evaluators for g, J_g, grad f and the Hessian of the
Lagrangian by algorithmic differentiation (reference optilayer.py:49-60).  Here
the problem is a polynomial in x with parameter-only coefficients, so all
derivatives are generated symbolically once, as term lists

    out[s] += coef * V[cidx] * prod_k x_ext[xi[k]]      (* lam_ext[lrow] for W)

grouped by output slot s (CSR).  ``V`` is the per-instance *parameter tape*:
V[0] = 1, V[1..n_par] = p, and further entries are parameter-only expressions
(products/sums, 1/p, indicators, sin/cos) evaluated level by level once per
solve.  ``x_ext`` is x with a trailing 1 so that padded factors are no-ops.

Row/variable order is the reference's flat order:
children in insertion order, entries in definition order, matrices column-major
(reference optilayer.py:225-272).
"""
from collections import OrderedDict

import numpy as np

from .poly import Poly, resolve, sym_info

FUNC_CODE = {'id': 0, 'inv': 1, 'ge': 2, 'gt': 3, 'sin': 4, 'cos': 5,
             'sqrt': 6}
MAX_FAC = 4          # V-factors per tape term (longer products are chained)
PRUNE = 1e-14
SPLIT_MAX = 4       # see lower(): coefficients of intermediates


class TermList(object):
    """CSR list of polynomial terms per output slot."""

    def __init__(self, n_out, width, with_lrow=False):
        self.n_out = n_out
        self.width = max(1, width)
        self.slots = [[] for _ in range(n_out)]
        self.with_lrow = with_lrow

    def add(self, slot, coef, cidx, xmon, lrow=0):
        self.slots[slot].append((coef, cidx, tuple(xmon), lrow))

    def finalize(self, n):
        nt = sum(len(s) for s in self.slots)
        self.ptr = np.zeros(self.n_out + 1, dtype=np.int32)
        self.coef = np.zeros(nt, dtype=np.float64)
        self.cidx = np.zeros(nt, dtype=np.int32)
        self.xi = np.full((nt, self.width), n, dtype=np.int32)
        self.lrow = np.zeros(nt, dtype=np.int32)
        k = 0
        for s, terms in enumerate(self.slots):
            # merge identical (cidx, xmon, lrow) terms
            merged = OrderedDict()
            for coef, cidx, xmon, lrow in terms:
                key = (cidx, xmon, lrow)
                merged[key] = merged.get(key, 0.0) + coef
            for (cidx, xmon, lrow), coef in merged.items():
                if coef == 0.0:
                    continue
                self.coef[k] = coef
                self.cidx[k] = cidx
                self.xi[k, :len(xmon)] = xmon
                self.lrow[k] = lrow
                k += 1
            self.ptr[s + 1] = k
        self.coef = self.coef[:k].copy()
        self.cidx = self.cidx[:k].copy()
        self.xi = np.ascontiguousarray(self.xi[:k])
        self.lrow = self.lrow[:k].copy()
        del self.slots
        return self

    @property
    def n_terms(self):
        return len(self.coef)


class _Tape(object):
    """Builder of the parameter tape V."""

    def __init__(self, par_index):
        self.par_index = par_index            # resolved sid -> k
        self.n_par = len(par_index)
        self.entries = []                     # (func, [(coef, [vidx...])], level)
        self.level = {0: 0}
        for k in range(self.n_par):
            self.level[1 + k] = 0
        self.sym_entry = {}                   # atom sid -> V idx
        self.poly_entry = {}                  # poly key -> V idx

    def _new_entry(self, func, terms):
        lvl = 1 + max([self.level[f] for _, fac in terms for f in fac] + [0])
        idx = 1 + self.n_par + len(self.entries)
        self.entries.append((func, terms, lvl))
        self.level[idx] = lvl
        return idx

    def v_of_symbol(self, sid):
        sid = resolve(sid)
        if sid in self.par_index:
            return 1 + self.par_index[sid]
        info = sym_info(sid)
        if info.kind != 'atom':
            raise ValueError('symbol %r is neither parameter nor atom: a '
                             'constraint coefficient depends on it' % info)
        if sid not in self.sym_entry:
            terms = self._terms_of(self._ppoly(info.arg))
            self.sym_entry[sid] = self._new_entry(info.func, terms)
        return self.sym_entry[sid]

    def _ppoly(self, poly):
        """Poly over parameter symbols -> {pmon(resolved sids): coef}."""
        out = {}
        for mono, c in poly.t.items():
            pm = tuple(sorted(resolve(s) for s in mono))
            out[pm] = out.get(pm, 0.0) + c
        return out

    def _product(self, vidx):
        vidx = sorted(vidx)
        while len(vidx) > MAX_FAC:
            head, vidx = vidx[:MAX_FAC], vidx[MAX_FAC:]
            key = ('prod', tuple(head))
            if key not in self.poly_entry:
                self.poly_entry[key] = self._new_entry('id', [(1.0, head)])
            vidx = sorted([self.poly_entry[key]] + vidx)
        return vidx

    def _terms_of(self, ppoly):
        terms = []
        for pm in sorted(ppoly):
            c = ppoly[pm]
            if c == 0.0:
                continue
            fac = self._product([self.v_of_symbol(s) for s in pm])
            terms.append((c, fac))
        return terms

    def v_of_ppoly(self, ppoly):
        """(scale, V index) such that ppoly == scale * V[index]."""
        ppoly = {pm: c for pm, c in ppoly.items() if abs(c) > PRUNE}
        if not ppoly:
            return 0.0, 0
        if len(ppoly) == 1:
            (pm, c), = ppoly.items()
            if len(pm) == 0:
                return c, 0
            if len(pm) == 1:
                return c, self.v_of_symbol(pm[0])
        lead_pm = sorted(ppoly)[0]
        scale = ppoly[lead_pm]
        key = tuple((pm, ppoly[pm] / scale) for pm in sorted(ppoly))
        if key not in self.poly_entry:
            terms = self._terms_of({pm: c for pm, c in key})
            self.poly_entry[key] = self._new_entry('id', terms)
        return scale, self.poly_entry[key]

    def finalize(self):
        """Sort entries by level; return arrays and the index remap."""
        base = 1 + self.n_par
        order = sorted(range(len(self.entries)),
                       key=lambda e: (self.entries[e][2], e))
        remap = np.arange(base + len(self.entries), dtype=np.int64)
        for new, old in enumerate(order):
            remap[base + old] = base + new
        func = np.zeros(len(order), dtype=np.int32)
        ptr = np.zeros(len(order) + 1, dtype=np.int32)
        coef, fac = [], []
        levels = []
        for new, old in enumerate(order):
            f, terms, lvl = self.entries[old]
            func[new] = FUNC_CODE[f]
            levels.append(lvl)
            for c, fs in terms:
                coef.append(c)
                row = [int(remap[v]) for v in fs] + [0] * (MAX_FAC - len(fs))
                fac.append(row)
            ptr[new + 1] = len(coef)
        n_levels = max(levels) if levels else 0
        level_ptr = np.zeros(n_levels + 1, dtype=np.int32)
        for lvl in levels:
            level_ptr[lvl] += 1
        level_ptr = np.concatenate([[0], np.cumsum(level_ptr[1:])]).astype(np.int32)
        return dict(n_v=base + len(order), tape_func=func, tape_ptr=ptr,
                    tape_coef=np.array(coef, dtype=np.float64),
                    tape_fac=np.array(fac, dtype=np.int32).reshape(-1, MAX_FAC),
                    level_ptr=level_ptr), remap


class NLPTables(object):
    """Flat description of  min f(x,p)  s.t.  lbg <= g(x,p) <= ubg."""

    def summary(self):
        return ('n=%d m=%d n_par=%d n_v=%d | terms g=%d J=%d(nnz %d) '
                'W=%d(nnz %d) f=%d | deg=%d' % (
                    self.n, self.m, self.n_par, self.n_v, self.G.n_terms,
                    self.J.n_terms, self.nnz_j, self.W.n_terms, self.nnz_w,
                    self.F.n_terms, self.degree))


def _split(poly, x_index, tape):
    """Poly -> {xmon: {pmon: coef}} with xmon sorted x indices."""
    out = {}
    for mono, c in poly.t.items():
        xm, pm = [], []
        for s in mono:
            r = resolve(s)
            j = x_index.get(r)
            if j is None:
                pm.append(r)
            else:
                xm.append(j)
        xm, pm = tuple(sorted(xm)), tuple(sorted(pm))
        d = out.setdefault(xm, {})
        d[pm] = d.get(pm, 0.0) + c
    return out


_FD_CACHE = {}


def _first_derivs(xmon):
    """d(prod x)/dx_j for every distinct j: [(j, mult, reduced monomial)]."""
    res = _FD_CACHE.get(xmon)
    if res is None:
        if len(_FD_CACHE) > 2000000:
            _FD_CACHE.clear()
        res = _FD_CACHE[xmon] = _first_derivs_uncached(xmon)
    return res


def _first_derivs_uncached(xmon):
    res = []
    for j in sorted(set(xmon)):
        mult = xmon.count(j)
        red = list(xmon)
        red.remove(j)
        res.append((j, float(mult), tuple(red)))
    return res


def lower(var_ids, par_ids, rows, objective, lbg, ubg, order_hint=None):
    """Build NLPTables.

    var_ids / par_ids : flat lists of (resolved) symbol ids defining x and p
    rows              : list of Poly (or float) constraint rows, flat order
    objective         : Poly
    """
    x_index = {resolve(s): j for j, s in enumerate(var_ids)}
    p_index = {resolve(s): k for k, s in enumerate(par_ids)}
    if len(x_index) != len(var_ids) or len(p_index) != len(par_ids):
        raise ValueError('duplicate symbols in variable/parameter lists')
    n, m = len(var_ids), len(rows)
    tape = _Tape(p_index)

    def as_poly(e):
        if isinstance(e, Poly):
            return e
        if isinstance(e, np.ndarray) and e.size == 1:
            return as_poly(e.reshape(-1)[0])
        c = float(e)
        return Poly({(): c} if c != 0.0 else {})

    rows = [as_poly(r) for r in rows]
    objective = as_poly(objective)
    # intermediates ('mid' symbols): x_ext = [x, 1, mids]; their definitions
    # become pseudo-rows m+1.. of the G/J/W term lists (slot m of lam_ext is the
    # objective factor) and the rows' derivatives follow by the chain rule
    mids = sorted({resolve(s) for r in rows for s in r.symbols()
                   if sym_info(resolve(s)).kind == 'mid'})
    n_mid = len(mids)
    for l, sid in enumerate(mids):
        x_index[sid] = n + 1 + l
    mid_defs = [sym_info(sid).arg for sid in mids]
    if any(sym_info(resolve(s)).kind == 'mid'
           for d in mid_defs + [objective] for s in d.symbols()):
        raise NotImplementedError('nested intermediates / intermediates in '
                                  'the objective are not supported')

    split_rows = [_split(r, x_index, tape) for r in rows + mid_defs]
    split_obj = _split(objective, x_index, tape)
    degree = max([len(xm) for sr in split_rows + [split_obj] for xm in sr] + [1])
    for sr in split_rows[:m]:
        for xm in sr:
            if sum(1 for j in xm if j > n) > 2:
                raise NotImplementedError(
                    'rows may contain at most two intermediates per monomial')

    def coef_of(ppoly):
        return tape.v_of_ppoly(ppoly)

    def lowered(sr):
        out = []
        for xm, pp in sorted(sr.items()):
            if len(xm) == 1 and xm[0] > n and 1 < len(pp) <= SPLIT_MAX:
                # coefficient of an intermediate: a few monomials over shared
                # parameter symbols, distinct for every (row, mid) pair -- one
                # term per monomial keeps the tape small (shared-memory sized)
                out += [(xm,) + coef_of({pm: c}) for pm, c in sorted(pp.items())]
            else:
                out.append((xm,) + coef_of(pp))
        return out

    # resolve all coefficients first (tape indices are remapped afterwards)
    lowered_rows = [lowered(sr) for sr in split_rows]
    lowered_obj = [(xm,) + coef_of(pp) for xm, pp in sorted(split_obj.items())]
    tape_arrays, remap = tape.finalize()

    G = TermList(m + n_mid, degree)
    F = TermList(1, degree)
    DF = TermList(n, degree - 1)
    jslots = OrderedDict()     # (row, col) -> list of (coef, cidx, xmon)
    wslots = OrderedDict()     # (j, k), j>=k -> list of (coef, cidx, xmon, lrow)

    def add_hess(xm, c, cidx, lrow):
        for j, mj, red in _first_derivs(xm):
            for k, mk, red2 in _first_derivs(red):
                if k > j:
                    continue
                # d2/dxj dxk of prod: mj*mk*(reduced); for j==k: mj*(mj-1)
                wslots.setdefault((j, k), []).append(
                    (c * mj * mk, cidx, red2, lrow))

    for i, terms in enumerate(lowered_rows):
        for xm, scale, cidx in terms:
            if scale == 0.0:
                continue
            cidx = int(remap[cidx])
            G.add(i, scale, cidx, xm)
            for j, mj, red in _first_derivs(xm):
                jslots.setdefault((i, j), []).append((scale * mj, cidx, red))
            if len(xm) >= 2:
                add_hess(xm, scale, cidx, i if i < m else i + 1)
    for xm, scale, cidx in lowered_obj:
        if scale == 0.0:
            continue
        cidx = int(remap[cidx])
        F.add(0, scale, cidx, xm)
        for j, mj, red in _first_derivs(xm):
            DF.add(j, scale * mj, cidx, red)
        if len(xm) >= 2:
            add_hess(xm, scale, cidx, m)

    # Jacobian slots: [real (row<m, col<n) incl. chain-rule fill | A = d row /
    # d mid | C = d mid / d x]
    akeys = sorted(k for k in jslots if k[0] < m and k[1] > n)
    ckeys = sorted(k for k in jslots if k[0] >= m)
    real = {k for k in jslots if k[0] < m and k[1] < n}
    c_by_mid = {}
    for (i, j) in ckeys:
        c_by_mid.setdefault(i - m, []).append(j)
    for (i, jm) in akeys:
        for j in c_by_mid.get(jm - n - 1, ()):
            real.add((i, j))
    jkeys = sorted(real)
    allkeys = jkeys + akeys + ckeys
    slot_of = {k: s for s, k in enumerate(allkeys)}
    J = TermList(len(allkeys), degree - 1)
    for s, key in enumerate(allkeys):
        for c, cidx, red in jslots.get(key, ()):
            J.add(s, c, cidx, red)
    # Hessian slots: [regular (j >= k, both variables) | cross (mid l, variable k) |
    # mid-mid (mid l1 >= mid l2)].  A cross slot holds X[l,k] = sum_i lam_i d2 row_i /
    # d mid_l d x_k (rows whose mid coefficient depends on x, e.g. hyperplane normal times
    # integrated position), a mid-mid slot M[l1,l2] = sum_i lam_i d2 row_i / d mid_l1 d
    # mid_l2 (rows with a product of two intermediates, e.g. the steering-rate rows of the
    # bicycle).  With C = d mid / d x the Hessian gains  X^T C + C^T X + C^T M C, gathered
    # per H position through the product lists xq_* below.
    wkeys = sorted(k for k in wslots if k[0] < n)
    xkeys = sorted(k for k in wslots if k[0] > n and k[1] < n)
    mkeys = sorted(k for k in wslots if k[0] > n and k[1] > n)
    W = TermList(len(wkeys) + len(xkeys) + len(mkeys), degree - 2, with_lrow=True)
    for s, key in enumerate(wkeys + xkeys + mkeys):
        for c, cidx, red, lrow in wslots[key]:
            W.add(s, c, cidx, red, lrow)

    tb = NLPTables()
    tb.n, tb.m, tb.n_par, tb.degree = n, m, len(par_ids), degree
    tb.n_mid = n_mid
    for k, v in tape_arrays.items():
        setattr(tb, k, v)
    tb.G, tb.F, tb.DF = G.finalize(n), F.finalize(n), DF.finalize(n)
    tb.J, tb.W = J.finalize(n), W.finalize(n)
    tb.jrow = np.array([k[0] for k in jkeys], dtype=np.int32)
    tb.jcol = np.array([k[1] for k in jkeys], dtype=np.int32)
    tb.jrow_ptr = np.searchsorted(tb.jrow, np.arange(m + 1)).astype(np.int32)
    tb.wrow = np.array([k[0] for k in wkeys], dtype=np.int32)
    tb.wcol = np.array([k[1] for k in wkeys], dtype=np.int32)
    tb.nnz_j, tb.nnz_w = len(jkeys), len(wkeys)
    tb.nnz_wx = len(xkeys) + len(mkeys)      # extra W slots (values Wx)
    tb.nnz_jx = len(allkeys)
    # products per lower-triangle H position (a >= b): Wx[w] * Jx[ca] * (Jx[cb] or 1)
    xpairs = {}
    for sx, (jm, k) in enumerate(xkeys):     # H[a,b] += X[l,k] C[l,j] (+ transpose)
        l = jm - n - 1
        for j in c_by_mid.get(l, ()):
            e = (sx, slot_of[(m + l, j)], -1)
            xpairs.setdefault((max(k, j), min(k, j)), []).append(e)
            if j == k:                        # both transposes land on the diagonal
                xpairs[(k, k)].append(e)
    for sm, (jm1, jm2) in enumerate(mkeys):  # H[a,b] += sum C[l1,a] M[l1,l2] C[l2,b]
        l1, l2 = jm1 - n - 1, jm2 - n - 1
        w = len(xkeys) + sm
        for la, lb in ((l1, l2),) + (((l2, l1),) if l1 != l2 else ()):
            for a in c_by_mid.get(la, ()):
                for b in c_by_mid.get(lb, ()):
                    if a >= b:
                        xpairs.setdefault((a, b), []).append(
                            (w, slot_of[(m + la, a)], slot_of[(m + lb, b)]))
    tb._xpairs = xpairs
    # chain rule  J[s] += sum_l A[i,l] C[l,j]  and  mu_l = sum_i lam_i A[i,l]
    a_by_row = {}
    for (i, jm) in akeys:
        a_by_row.setdefault(i, []).append((jm - n - 1, slot_of[(i, jm)]))
    jp_ptr, jp_a, jp_c = [0], [], []
    for (i, j) in jkeys:
        for l, sa in a_by_row.get(i, ()):
            sc = slot_of.get((m + l, j))
            if sc is not None:
                jp_a.append(sa)
                jp_c.append(sc)
        jp_ptr.append(len(jp_a))
    tb.jp_ptr = np.array(jp_ptr, dtype=np.int32)
    tb.jp_a = np.array(jp_a, dtype=np.int32)
    tb.jp_c = np.array(jp_c, dtype=np.int32)
    mu_ptr, mu_row, mu_slot = [0], [], []
    by_mid = {}
    for (i, jm) in akeys:
        by_mid.setdefault(jm - n - 1, []).append((i, slot_of[(i, jm)]))
    for l in range(n_mid):
        for i, sa in by_mid.get(l, ()):
            mu_row.append(i)
            mu_slot.append(sa)
        mu_ptr.append(len(mu_row))
    tb.mu_ptr = np.array(mu_ptr, dtype=np.int32)
    tb.mu_row = np.array(mu_row, dtype=np.int32)
    tb.mu_slot = np.array(mu_slot, dtype=np.int32)
    tb.lbg = np.asarray(lbg, dtype=np.float64).copy()
    tb.ubg = np.asarray(ubg, dtype=np.float64).copy()
    _build_kkt_pattern(tb)
    build_kkt_structure(tb, order_hint)
    return tb


def _build_kkt_pattern(tb):
    """Gather lists for the condensed KKT matrix  H = W + J^T Sigma J.

    For every structurally non-zero lower-triangular position (j >= k) the
    contributing pairs of Jacobian slots (s1, s2) of a common row: a
    deterministic gather instead of scattered atomics.
    """
    pos = {}
    for s in range(tb.nnz_w):
        pos.setdefault((int(tb.wrow[s]), int(tb.wcol[s])), [])
    xpairs = getattr(tb, '_xpairs', {})
    for key in xpairs:
        pos.setdefault(key, [])
    for i in range(tb.m):
        lo, hi = int(tb.jrow_ptr[i]), int(tb.jrow_ptr[i + 1])
        for s1 in range(lo, hi):
            for s2 in range(lo, s1 + 1):
                j, k = int(tb.jcol[s1]), int(tb.jcol[s2])   # j >= k (sorted)
                pos.setdefault((j, k), []).append((s1, s2, i))
    keys = sorted(pos)
    tb.hrow = np.array([k[0] for k in keys], dtype=np.int32)
    tb.hcol = np.array([k[1] for k in keys], dtype=np.int32)
    tb.nnz_h = len(keys)
    ptr = np.zeros(len(keys) + 1, dtype=np.int32)
    s1l, s2l, rowl = [], [], []
    for q, key in enumerate(keys):
        for s1, s2, i in pos[key]:
            s1l.append(s1)
            s2l.append(s2)
            rowl.append(i)
        ptr[q + 1] = len(s1l)
    tb.hp_ptr = ptr
    tb.hp_s1 = np.array(s1l, dtype=np.int32)
    tb.hp_s2 = np.array(s2l, dtype=np.int32)
    tb.hp_row = np.array(rowl, dtype=np.int32)
    # position of every W slot inside the H pattern
    index = {k: q for q, k in enumerate(keys)}
    tb.w2h = np.array([index[(int(r), int(c))]
                       for r, c in zip(tb.wrow, tb.wcol)], dtype=np.int32)
    # gather lists of the extra Hessian products: for the H positions xq_h[e] the entries
    # [xq_ptr[e], xq_ptr[e+1]) of (extra W slot xq_w, Jacobian slots xq_a, xq_b of C;
    # xq_b = -1: no second factor)
    xq_h, xq_ptr, xq_w, xq_a, xq_b = [], [0], [], [], []
    for key in sorted(xpairs):
        xq_h.append(index[key])
        for sx, sa, sb in xpairs[key]:
            xq_w.append(sx)
            xq_a.append(sa)
            xq_b.append(sb)
        xq_ptr.append(len(xq_w))
    tb.n_xq = len(xq_h)
    tb.xq_h = np.array(xq_h, dtype=np.int32)
    tb.xq_ptr = np.array(xq_ptr, dtype=np.int32)
    tb.xq_w = np.array(xq_w, dtype=np.int32)
    tb.xq_a = np.array(xq_a, dtype=np.int32)
    tb.xq_b = np.array(xq_b, dtype=np.int32)
    if hasattr(tb, '_xpairs'):
        del tb._xpairs


# ---------------------------------------------------------------------------
# KKT ordering + envelope (skyline) structure
# ---------------------------------------------------------------------------
KKT_NB = 8       # panel width of the blocked factorisation (csrc/omg_b200.cu)


def _envelope_first(adj_lower_rows, perm_pos, N):
    """first[i] (permuted) = smallest permuted column index coupled to row i."""
    first = np.arange(N)
    for a, cols in adj_lower_rows.items():
        pa = perm_pos[a]
        for b in cols:
            pb = perm_pos[b]
            lo, hi = (pb, pa) if pb < pa else (pa, pb)
            if lo < first[hi]:
                first[hi] = lo
    return first


def build_kkt_structure(tb, hint=None):
    """Choose a symmetric permutation of the condensed KKT matrix

        K = [[H, Jc^T], [Jc, -delta_c I]],   H = W + J^T Sigma J  (n x n)

    and lay out its lower envelope.  B-spline locality makes K banded once the
    unknowns are ordered by "time"; the envelope of the factor L (K = L S L^T,
    S = diag(+1 variables, -1 equality rows)) equals the envelope of K, so the
    factorisation needs ~7x fewer flops and ~2.5x less storage than dense
    packed storage.  Each equality row is placed right after the last variable
    it couples, so its pivot is a (negative) Schur complement.

    hint: optional array [n] of relative positions in [0,1] (spline
    coefficient index / length) used as the ordering key; a reverse
    Cuthill-McKee ordering of the pattern is the alternative, the one with the
    smaller envelope wins.
    """
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    n = tb.n
    eq_rows = np.nonzero(tb.lbg == tb.ubg)[0].astype(np.int32)
    n_eq = len(eq_rows)
    N = n + n_eq
    # adjacency (lower part, natural numbering: variables then equality rows)
    adj = {}
    for r, c in zip(tb.hrow, tb.hcol):
        adj.setdefault(int(r), set()).add(int(c))
    eq_cols = []
    for k, i in enumerate(eq_rows):
        cols = [int(tb.jcol[s]) for s in range(tb.jrow_ptr[i], tb.jrow_ptr[i + 1])]
        eq_cols.append(cols)
        adj.setdefault(n + k, set()).update(cols)

    def finish(order_vars_key):
        """order by key, then insert equality rows after their last variable."""
        order = list(np.argsort(order_vars_key, kind='stable'))
        rank = {v: r for r, v in enumerate(order)}
        ins = {}
        for k, cols in enumerate(eq_cols):
            last = max(rank[c] for c in cols)
            ins.setdefault(last, []).append(n + k)
        seq = []
        for r, v in enumerate(order):
            seq.append(v)
            seq += ins.get(r, [])
        pos = np.empty(N, dtype=np.int64)
        pos[np.array(seq)] = np.arange(N)
        first = _envelope_first(adj, pos, N)
        first = (first // KKT_NB) * KKT_NB
        # multiply-adds of the envelope factorisation: entry (i, j) costs the overlap of the two
        # rows' envelopes left of column j
        flops = 0
        for i in range(N):
            js = np.arange(first[i], i + 1)
            flops += int(np.maximum(0, js - np.maximum(first[i], first[js])).sum())
        return pos, first, int(np.sum(np.arange(N) - first + 1)), flops

    cands = []
    rows = np.concatenate([tb.hrow, tb.hcol]).astype(np.int64)
    cols = np.concatenate([tb.hcol, tb.hrow]).astype(np.int64)
    A = coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)).tocsr()
    rcm = np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True))
    key = np.empty(n)
    key[rcm] = np.arange(n)
    cands.append(finish(key))
    cands.append(finish(-key))
    if hint is not None:
        cands.append(finish(np.asarray(hint, dtype=float)))
    cands.append(finish(np.arange(n, dtype=float)))
    # "arrow" orderings: the unknowns coupled to (nearly) everything -- e.g. the slack splines of
    # the quadrotor's acceleration rows -- go LAST as a dense border, the rest is ordered by RCM of
    # its own sub-graph and stays narrowly banded.  (Quadrotor3D with 5 plates: 68 k envelope
    # entries / 8.1 M multiply-adds with the plain orderings, 20 k / 0.5 M with the border.)
    arrows = []
    deg = np.asarray(A.getnnz(axis=1)).ravel()
    base = cands[0][0][:n].astype(float) / max(N, 1)
    for thr in sorted(set(int(d) for d in deg if d > np.median(deg)))[:8]:
        low = np.nonzero(deg < thr)[0]
        if len(low) < 2 or len(low) == n:
            continue
        sub = reverse_cuthill_mckee(A[low][:, low].tocsr(), symmetric_mode=True)
        key = base + 10.0
        key[low[np.asarray(sub)]] = np.arange(len(low)) / float(len(low))
        arrows.append(finish(key))
    pos, first, size, flops = min(cands, key=lambda c: c[2])
    if arrows:      # adopted only where the factorisation is heavy and the border at least halves it
        # (small problems keep their band ordering: nothing to gain, and their long cold starts
        #  -- Dubins, ~250 iterations -- are sensitive to any change of the summation order)
        best = min(arrows, key=lambda c: (c[3], c[2]))
        if flops >= 1000000 and 2 * best[3] <= flops:
            pos, first, size, flops = best

    tb.kkt_n = N
    tb.kkt_n_eq = n_eq
    tb.kkt_eq_rows = eq_rows
    tb.kkt_pos_var = pos[:n].astype(np.int32)
    tb.kkt_pos_eq = pos[n:].astype(np.int32)
    sign = np.ones(N, dtype=np.int32)
    sign[tb.kkt_pos_eq] = -1
    tb.kkt_sign = sign
    # rows 0..N-1 of K plus the right-hand-side row N (dense)
    first_all = np.concatenate([first, [0]]).astype(np.int32)
    width = np.arange(N + 1) - first_all + 1
    width[N] = N                         # rhs row has no diagonal entry
    ptr = np.concatenate([[0], np.cumsum(width)]).astype(np.int32)
    tb.env_first, tb.env_ptr = first_all, ptr
    tb.env_size = int(ptr[-1])            # natural (unpadded) row widths spread the
    # rows of a panel over the shared-memory banks

    def env(i, j):
        return int(ptr[i] + (j - first_all[i]))

    # destinations of the assembly gathers
    hdst = np.empty(tb.nnz_h, dtype=np.int32)
    for q, (r, c) in enumerate(zip(tb.hrow, tb.hcol)):
        pr, pc = pos[r], pos[c]
        hi, lo = (pr, pc) if pr >= pc else (pc, pr)
        hdst[q] = env(hi, lo)
    tb.kkt_hdst = hdst
    jdst = np.full(tb.nnz_j, -1, dtype=np.int32)
    for k, i in enumerate(eq_rows):
        pk = pos[n + k]
        for s in range(tb.jrow_ptr[i], tb.jrow_ptr[i + 1]):
            pc = pos[tb.jcol[s]]
            hi, lo = (pk, pc) if pk >= pc else (pc, pk)
            jdst[s] = env(hi, lo)
    tb.kkt_jdst = jdst
    tb.kkt_diag = np.array([env(i, i) for i in range(N)], dtype=np.int32)
    # panels and the rows each panel reaches
    n_pan = (N + KKT_NB - 1) // KKT_NB
    prow_ptr = [0]
    prows = []
    for pb in range(n_pan):
        kb = pb * KKT_NB
        ke = min(N, kb + KKT_NB)
        rr = [i for i in range(ke, N + 1) if first_all[i] < ke]
        prows += rr
        prow_ptr.append(len(prows))
    tb.kkt_panel_ptr = np.array(prow_ptr, dtype=np.int32)
    tb.kkt_panel_rows = np.array(prows, dtype=np.int32)
    tb.kkt_max_panel_rows = int(np.max(np.diff(prow_ptr))) if n_pan else 0
    return tb
