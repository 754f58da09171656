"""B-spline bases and coefficient-space spline algebra.

Host-side model-building code (runs once per problem structure).  It mirrors
the semantics of the reference's ``omgtools/basics/spline.py`` so that every
constraint block has the same rows in the same order:

* knot-union rule for sums/products            -- spline.py:138-148
* indicator conventions of Cox-de Boor         -- spline.py:131-136, 214-233
* derivative matrices                          -- spline.py:236-260
* basis change by collocation at the arg-max of each basis function on a
  501-point grid, entries below 1e-10 dropped  -- spline.py:36, 280-306
* product = transform of pairwise coefficient products -- spline.py:419-436

Coefficients may be floats or ``Poly`` objects (symbolic variables /
parameters); all matrices are plain numpy.  The matrices produced here end up,
through ``lowering.py``, as the constant tables of the CUDA kernels.
"""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/basics/spline.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
from collections import Counter

import numpy as np
import scipy.linalg as la

from .poly import Poly, matvec, is_symbolic

NO_POINTS = 501
_DROP_TOL = 1e-10

_BASIS_CACHE = {}
_PRODUCT_CACHE = {}
_TRANSFORM_CACHE = {}


def _dot(T, coeffs):
    coeffs = np.asarray(coeffs)
    if coeffs.dtype == object:
        return matvec(T, coeffs)
    return np.asarray(T).dot(coeffs)


class BSplineBasis(object):
    """Numerical B-spline basis (knots, degree); instances are interned."""

    def __new__(cls, knots, degree):
        knots = np.array(knots, dtype=float)
        key = (knots.tobytes(), int(degree))
        inst = _BASIS_CACHE.get(key)
        if inst is None:
            inst = object.__new__(cls)
            inst.knots = knots
            inst.degree = int(degree)
            inst._key = key
            inst._x = np.linspace(knots[0], knots[-1], NO_POINTS)
            inst._grid_eval = None
            _BASIS_CACHE[key] = inst
        return inst

    def __len__(self):
        return len(self.knots) - self.degree - 1

    def __eq__(self, other):
        return isinstance(other, BSplineBasis) and self._key == other._key

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._key)

    def __repr__(self):
        return 'BSplineBasis(deg=%d, n=%d)' % (self.degree, len(self))

    # ---- evaluation ------------------------------------------------------
    def eval_basis(self, x):
        """Matrix B[i, l] = l-th basis function at x[i] (Cox-de Boor).

        Interval convention (reference spline.py:131-136): the leading
        clamped interval is closed on both sides, all others are (k_i, k_i+1].
        """
        x = np.atleast_1d(np.asarray(x, dtype=float))
        k, p = self.knots, self.degree
        n_int = len(k) - 1
        lvl = np.zeros((len(x), n_int))
        for i in range(n_int):
            if i < p + 1 and k[0] == k[i]:
                lvl[:, i] = (x >= k[i]) & (x <= k[i + 1])
            else:
                lvl[:, i] = (x > k[i]) & (x <= k[i + 1])
        for d in range(1, p + 1):
            nxt = np.zeros((len(x), n_int - d))
            for i in range(n_int - d):
                den = k[i + d] - k[i]
                if den != 0:
                    nxt[:, i] += (x - k[i]) * lvl[:, i] / den
                den = k[i + d + 1] - k[i + 1]
                if den != 0:
                    nxt[:, i] += (k[i + d + 1] - x) * lvl[:, i + 1] / den
            lvl = nxt
        return lvl

    __call__ = eval_basis

    def _on_grid(self):
        if self._grid_eval is None:
            self._grid_eval = self.eval_basis(self._x)
        return self._grid_eval

    def greville(self):
        p = self.degree
        return [sum(self.knots[i + 1:i + p + 1]) / float(p)
                for i in range(len(self))]

    def support(self):
        p = self.degree
        return list(zip(self.knots[:-(p + 1)], self.knots[p + 1:]))

    # ---- algebra on bases ----------------------------------------------------
    def _combine(self, other, degree):
        """Knot union with multiplicity max(m1+degree-p1, m2+degree-p2)
        (reference spline.py:138-148)."""
        c1, c2 = Counter(self.knots.tolist()), Counter(other.knots.tolist())
        knots = []
        for b in sorted(set(c1) | set(c2)):
            m = max(c1.get(b, -np.inf) + degree - self.degree,
                    c2.get(b, -np.inf) + degree - other.degree)
            knots += [b] * int(m)
        return BSplineBasis(knots, degree)

    def __add__(self, other):
        if isinstance(other, BSplineBasis):
            return self._combine(other, max(self.degree, other.degree))
        if isinstance(other, (int, float)):
            return self
        raise TypeError('cannot add %r to a basis' % (other,))

    __radd__ = __add__
    __sub__ = __add__
    __rsub__ = __add__

    def __mul__(self, other):
        if isinstance(other, BSplineBasis):
            return self._combine(other, self.degree + other.degree)
        if isinstance(other, (int, float)):
            return self
        raise TypeError('cannot multiply a basis with %r' % (other,))

    __rmul__ = __mul__

    def __pow__(self, power):
        if not isinstance(power, int):
            raise TypeError('power must be integer')
        return self._combine(self, power * self.degree)

    def insert_knots(self, knots):
        extra = np.setdiff1d(knots, self.knots)
        return BSplineBasis(np.sort(np.append(self.knots, extra)), self.degree)

    def scale(self, factor, shift=0):
        return BSplineBasis(self.knots * factor + shift, self.degree)

    # ---- matrices ------------------------------------------------------------
    def derivative(self, o=1):
        """(basis of the o-th derivative, matrix P) with c' = P c
        (de Boor ch. X eq. 16; reference spline.py:236-260)."""
        p, N = self.degree, len(self)
        B = BSplineBasis(self.knots[o:-o], p - o)
        P = np.eye(N)
        knots = self.knots
        for i in range(o):
            knots = knots[1:-1]
            delta = knots[p - i:] - knots[:-(p - i)]
            rows = N - 1 - i
            D = np.zeros((rows, rows + 1))
            j = np.arange(rows)
            D[j, j] = -1. / delta
            D[j, j + 1] = 1. / delta
            P = (p - i) * D.dot(P)
        return B, P

    def pairs(self, other):
        """Index pairs (i, j) of basis functions with overlapping support."""
        ia, ib = [], []
        for i, (a0, a1) in enumerate(self.support()):
            for j, (b0, b1) in enumerate(other.support()):
                if max(a0, b0) < min(a1, b1):
                    ia.append(i)
                    ib.append(j)
        return np.array(ia), np.array(ib)

    def _collocation(self):
        b = self._on_grid()
        m = np.argmax(b, axis=0)
        return m, b[m, :]

    def transform(self, other):
        """T with self(x) T = other(x); ``other`` a basis or a callable that
        maps grid indices to function values (reference spline.py:280-306)."""
        if isinstance(other, BSplineBasis):
            if other == self:
                return np.eye(len(self))
            key = (self._key, other._key)
            T = _TRANSFORM_CACHE.get(key)
            if T is None:
                m, bm = self._collocation()
                T = la.solve(bm, other.eval_basis(self._x[m]))
                T[abs(T) < _DROP_TOL] = 0.
                _TRANSFORM_CACHE[key] = T
            return T
        m, bm = self._collocation()
        T = la.solve(bm, other(m))
        T[abs(T) < _DROP_TOL] = 0.
        return T


def product_transform(b1, b2):
    """(product basis, pairs, T) with coeffs(s1*s2) = T (c1[pairs0]*c2[pairs1])
    (reference spline.py:419-436)."""
    key = (b1._key, b2._key)
    hit = _PRODUCT_CACHE.get(key)
    if hit is None:
        basis = b1 * b2
        p0, p1 = b1.pairs(b2)
        e1 = b1.eval_basis(basis._x)
        e2 = b2.eval_basis(basis._x)
        prod = e1[:, p0] * e2[:, p1]
        T = basis.transform(lambda idx: prod[idx, :])
        hit = (basis, (p0, p1), T)
        _PRODUCT_CACHE[key] = hit
    return hit


class BSpline(object):
    """Spline = basis + coefficient vector (floats or Poly)."""

    def __init__(self, basis, coeffs):
        self.basis = basis
        self.coeffs = np.asarray(coeffs) if not isinstance(coeffs, np.ndarray) \
            else coeffs

    def __len__(self):
        return len(self.basis)

    def __call__(self, x):
        return _dot(self.basis.eval_basis(x), self.coeffs)

    # ---- vector-space operations ----------------------------------------
    def __add__(self, other):
        if isinstance(other, BSpline):
            basis = self.basis + other.basis
            return BSpline(basis,
                           _dot(basis.transform(self.basis), self.coeffs) +
                           _dot(basis.transform(other.basis), other.coeffs))
        return BSpline(self.basis, self.coeffs + other)

    __radd__ = __add__

    def __neg__(self):
        return BSpline(self.basis, -self.coeffs)

    def __sub__(self, other):
        return self + (-other)

    def __rsub__(self, other):
        return other + (-self)

    def __mul__(self, other):
        if isinstance(other, BSpline):
            basis, (p0, p1), T = product_transform(self.basis, other.basis)
            cp = self.coeffs[p0] * other.coeffs[p1]
            return BSpline(basis, _dot(T, cp))
        return BSpline(self.basis, other * self.coeffs)

    __rmul__ = __mul__

    def __pow__(self, power):
        if not isinstance(power, int):
            raise TypeError('exponent must be integer')
        res = self
        for _ in range(1, power):
            res = res * self
        return res

    # ---- calculus ----------------------------------------------------------
    def derivative(self, o=1):
        if o == 0:
            return self
        Bd, Pd = self.basis.derivative(o)
        return BSpline(Bd, _dot(Pd, self.coeffs))

    def integral(self):
        """Integral over the whole support (de Boor X.33,
        reference spline.py:478-487)."""
        k, d = self.basis.knots, self.basis.degree
        w = (k[d + 1:] - k[:-(d + 1)]) / (d + 1)
        total = 0.
        for wi, ci in zip(w, self.coeffs):
            total = total + ci * wi
        return total

    def insert_knots(self, knots):
        basis = self.basis.insert_knots(knots)
        return BSpline(basis, _dot(basis.transform(self.basis), self.coeffs))

    def scale(self, factor, shift=0):
        return BSpline(self.basis.scale(factor, shift), self.coeffs)
