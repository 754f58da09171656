"""Spline utilities used by the MPC hot path: point evaluation with a symbolic
abscissa, running/definite integrals, receding-horizon shift matrices and the
first-knot shift used by ADMM.

Semantics follow the reference's ``omgtools/basics/spline_extra.py`` (line
numbers in each docstring); the code is written for this framework's ``Poly``
scalars instead of CasADi ``MX``.
"""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/basics/spline_extra.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np
from scipy.interpolate import splev

from .poly import Poly, is_symbolic, as_object_vector
from .spline import BSpline, BSplineBasis, _dot


def _scalar(x):
    if isinstance(x, np.ndarray) and x.size == 1:
        return x.reshape(-1)[0]
    return x


def evalspline(s, x, share_weights=False):
    """Value of spline ``s`` at abscissa ``x`` (float or parameter Poly).

    Cox-de Boor recursion with indicator functions of x, so the result stays
    a polynomial in the coefficients with parameter-only weights
    (reference spline_extra.py:28-55).  ``share_weights`` names every basis
    value B_l(x) once (identity atom) instead of expanding it into each
    coefficient's monomials -- for long splines with symbolic coefficients.
    """
    x = _scalar(x)
    basis = s.basis
    k, p = basis.knots, basis.degree
    lvl = []
    for i in range(len(k) - 1):
        if i < p + 1 and k[0] == k[i]:
            lvl.append((x >= k[i]) * (x <= k[i + 1]))
        else:
            lvl.append((x > k[i]) * (x <= k[i + 1]))
    for d in range(1, p + 1):
        nxt = []
        for i in range(len(k) - d - 1):
            b = 0. * x
            den = k[i + d] - k[i]
            if den != 0:
                b = (x - k[i]) * lvl[i] / den
            den = k[i + d + 1] - k[i + 1]
            if den != 0:
                b = b + (k[i + d + 1] - x) * lvl[i + 1] / den
            nxt.append(b)
        lvl = nxt
    # numeric abscissa: plain floats (numpy bools / scalars do not combine with Poly)
    lvl = [b if isinstance(b, Poly) else float(b) for b in lvl]
    if share_weights:
        from .poly import share
        lvl = [share(b) if isinstance(b, Poly) else b for b in lvl]
    result = 0.
    for l in range(len(basis)):
        result = result + s.coeffs[l] * lvl[l]
    return result


def running_integral(spline):
    """Antiderivative spline (degree+1, one extra knot each side)
    (reference spline_extra.py:58-76)."""
    basis, coeffs = spline.basis, spline.coeffs
    knots, degree = basis.knots, basis.degree
    basis_int = BSplineBasis(np.r_[knots[0], knots, knots[-1]], degree + 1)
    acc = [0.]
    for i in range(len(basis_int) - 1):
        acc.append(acc[i] + (knots[degree + i + 1] - knots[i]) /
                   float(degree + 1) * coeffs[i])
    if is_symbolic(coeffs):
        cint = as_object_vector(np.array(acc, dtype=object))
    else:
        cint = np.array(acc, dtype=float)
    return BSpline(basis_int, cint)


def definite_integral(spline, a, b):
    """Integral over [a, b] (reference spline_extra.py:79-85)."""
    spline_int = running_integral(spline)
    return evalspline(spline_int, b) - evalspline(spline_int, a)


def shift_spline(coeffs, t_shift, basis):
    """Piece [t_shift, end] re-expressed on an equidistant basis
    (reference spline_extra.py:88-99)."""
    n_knots = len(basis) - basis.degree + 1
    knots, degree = basis.knots, basis.degree
    knots2 = np.r_[t_shift * np.ones(degree),
                   np.linspace(t_shift, knots[-1], n_knots),
                   knots[-1] * np.ones(degree)]
    basis2 = BSplineBasis(knots2, degree)
    return _dot(basis2.transform(basis), coeffs)


def _interior_multiplicity(knots, deg):
    m = 1
    while knots[-deg - 2 - m] >= knots[-deg - 2]:
        m += 1
    return m


def extrapolate_T(basis, t_extra):
    """(N+m) x N matrix extending the spline by one knot interval of length
    t_extra with maximal smoothness (reference spline_extra.py:107-155)."""
    knots, deg, N = basis.knots, basis.degree, len(basis)
    m = _interior_multiplicity(knots, deg)
    knots2 = np.r_[knots[:-deg - 1], knots[-deg - 1] * np.ones(m),
                   (knots[-1] + t_extra) * np.ones(deg + 1)]
    basis2 = BSplineBasis(knots2, deg)
    A = np.zeros((deg + 1, deg + 1))
    B = np.zeros((deg + 1, deg + 1))
    if m < deg + 1:
        # value conditions on the last (deg+1-m) Greville points
        pts = basis.greville()[-(deg + 1 - m):]
        a = basis2.eval_basis(pts)[:, -(deg + 1 + m):-m]
        b = basis.eval_basis(pts)[:, -(deg + 1):]
        A[:(deg + 1 - m), -(deg + 1):-m] = a[:, m:]
        B[:(deg + 1 - m), :m] = b[:, :m] - a[:, :m]
        B[:(deg + 1 - m), m:] = b[:, m:]
    else:
        A[0, -(deg + 1)] = 1.
        B[0, -1] = 1.
    # continuity of the m highest derivatives at the old end point
    A1, B1 = np.identity(deg + 1), np.identity(deg + 1)
    for i in range(1, deg + 1):
        q = deg + 1 - i
        Ad, Bd = np.zeros((q, q + 1)), np.zeros((q, q + 1))
        for j in range(q):
            wb = q / (knots[j + N] - knots[j + N - deg - 1 + i])
            wa = q / (knots2[j + N + m] - knots2[j + N - deg - 1 + m + i])
            Bd[j, j], Bd[j, j + 1] = -wb, wb
            Ad[j, j], Ad[j, j + 1] = -wa, wa
        A1, B1 = Ad.dot(A1), Bd.dot(B1)
        if i >= deg + 1 - m:
            A[i, :] = A1[-(deg - i + 1), :]
            B[i, :] = B1[-1, :]
    blk = np.linalg.solve(A, B)
    blk[abs(blk) < 1e-10] = 0.
    T = np.zeros((N + m, N))
    T[:N, :N] = np.eye(N)
    T[-(deg + 1):, -(deg + 1):] = blk
    return T


def shiftoverknot_T(basis):
    """N x N warm-start matrix: move the horizon one knot interval ahead and
    extrapolate the tail (reference spline_extra.py:165-191)."""
    knots, deg = basis.knots, basis.degree
    m = _interior_multiplicity(knots, deg)
    t_shift = knots[deg + 1] - knots[0]
    T = np.diag(np.ones(len(basis) - m), m)
    blk = np.eye(deg + 1)
    for k in range(deg):
        step = np.zeros((deg + 1 + k + 1, deg + 1 + k))
        for j in range(deg + 1 + k + 1):
            if j >= deg + 1:
                step[j, j - 1] = 1.
            elif j <= k:
                step[j, j] = 1.
            else:
                den = knots[j + deg - k] - knots[j]
                step[j, j - 1] = (knots[j + deg - k] - t_shift) / den
                step[j, j] = (t_shift - knots[j]) / den
        blk = step.dot(blk)
    T[:deg, :deg + 1] = blk[deg + 1:, :]
    T_extr = extrapolate_T(basis, knots[-1] - knots[-deg - 2])
    T[-(deg + 1):, -(deg + 1):] = T_extr[-(deg + 1):, -(deg + 1):]
    return T


def shift_over_knot(coeffs, basis):
    return _dot(shiftoverknot_T(basis), coeffs)


def shiftfirstknot_T(basis, t_shift, inverse=False):
    """Matrix restricting a spline to [t_shift, end] by moving its first
    degree+1 knots; t_shift may be a parameter Poly.  With inverse=True also
    the (upper-triangular block) inverse (reference spline_extra.py:220-255).
    Returns object matrices when t_shift is symbolic.
    """
    t_shift = _scalar(t_shift)
    knots, deg, N = basis.knots, basis.degree, len(basis)
    sym = isinstance(t_shift, Poly)

    def zeros(r, c):
        if sym:
            Z = np.empty((r, c), dtype=object)
            Z[:, :] = 0.
            return Z
        return np.zeros((r, c))

    def eye(n):
        Z = zeros(n, n)
        for i in range(n):
            Z[i, i] = 1.
        return Z

    def matmul(A, B):
        if not sym:
            return A.dot(B)
        C = zeros(A.shape[0], B.shape[1])
        for i in range(A.shape[0]):
            for j in range(B.shape[1]):
                acc = 0.
                for k in range(A.shape[1]):
                    a, b = A[i, k], B[k, j]
                    if (isinstance(a, float) and a == 0.) or \
                            (isinstance(b, float) and b == 0.):
                        continue
                    acc = acc + a * b
                C[i, j] = acc
        return C

    blk = eye(deg + 1)
    for k in range(deg + 1):
        step = zeros(deg + 1 + k + 1, deg + 1 + k)
        for j in range(deg + 1 + k + 1):
            if j >= deg + 1:
                step[j, j - 1] = 1.
            elif j <= k:
                step[j, j] = 1.
            else:
                den = knots[j + deg - k] - knots[j]
                step[j, j - 1] = (knots[j + deg - k] - t_shift) / den
                step[j, j] = (t_shift - knots[j]) / den
        blk = matmul(step, blk)
    T = eye(N)
    T[:deg + 1, :deg + 1] = blk[deg + 1:, :]
    if not inverse:
        return T
    Tinv = eye(N)
    for i in range(deg, -1, -1):
        Tinv[i, i] = 1. / T[i, i]
        for j in range(deg, i, -1):
            acc = 0.
            for k in range(i + 1, deg + 2):
                if k < N:
                    acc = acc + T[i, k] * Tinv[k, j]
            Tinv[i, j] = (-1. / T[i, i]) * acc
    return T, Tinv


def _symmat_dot(T, cfs):
    cfs = np.asarray(cfs)
    out = np.empty(T.shape[0], dtype=object)
    for i in range(T.shape[0]):
        acc = 0.
        for j in range(T.shape[1]):
            a = T[i, j]
            if isinstance(a, float) and a == 0.:
                continue
            acc = acc + a * cfs[j]
        out[i] = acc
    return out


def shift_knot1_fwd(cfs, basis, t_shift):
    """Coefficients of the spline restricted to [t_shift, end]
    (reference spline_extra.py:194-204)."""
    T = shiftfirstknot_T(basis, t_shift)
    if T.dtype == object or is_symbolic(cfs):
        return _symmat_dot(T, cfs)
    return T.dot(cfs)


def shift_knot1_bwd(cfs, basis, t_shift):
    """Inverse of shift_knot1_fwd (reference spline_extra.py:207-217)."""
    _, Tinv = shiftfirstknot_T(basis, t_shift, inverse=True)
    if Tinv.dtype == object or is_symbolic(cfs):
        return _symmat_dot(Tinv, cfs)
    return Tinv.dot(cfs)


def knot_insertion_T(basis, knots_to_insert):
    """Boehm knot insertion matrix (reference spline_extra.py:258-280)."""
    N = len(basis)
    knots = basis.knots.tolist()
    degree = basis.degree
    T = np.eye(N)
    for knot in knots_to_insert:
        step = np.zeros((N + 1, N))
        for j in range(N + 1):
            if knot <= knots[j]:
                w = 0.
            elif knots[j] < knot and knot < knots[j + degree]:
                w = (knot - knots[j]) / (knots[j + degree] - knots[j])
            else:
                w = 1.
            if j != 0:
                step[j, j - 1] = 1. - w
            if j != N:
                step[j, j] = w
        T = step.dot(T)
        N += 1
        knots = sorted(knots + [knot])
    return T, knots


def get_interval_T(basis, min_value, max_value):
    """Matrix extracting the piece [min_value, max_value]
    (reference spline_extra.py:283-295)."""
    knots, degree = basis.knots, basis.degree
    n_min = len(np.where(knots == min_value)[0])
    n_max = len(np.where(knots == max_value)[0])
    extra = [min_value] * (degree + 1 - n_min) + \
        [max_value] * (degree + 1 - n_max)
    T, knots2 = knot_insertion_T(basis, extra)
    jmin = np.searchsorted(knots2, min_value, side='left')
    jmax = np.searchsorted(knots2, max_value, side='right')
    return T[jmin:jmax - degree - 1, :], knots2[jmin:jmax]


def crop_spline(spline, min_value, max_value):
    T, knots2 = get_interval_T(spline.basis, min_value, max_value)
    return BSpline(BSplineBasis(knots2, spline.basis.degree),
                   _dot(T, spline.coeffs))


def sample_splines(spline, time):
    """Numeric sampling (reference spline_extra.py:406-410)."""
    if isinstance(spline, list):
        return [sample_splines(s, time) for s in spline]
    return splev(time, (spline.basis.knots,
                        np.asarray(spline.coeffs, dtype=float),
                        spline.basis.degree))
