"""Spline algebra vs golden matrices produced by the REFERENCE's spline.py /
spline_extra.py (tests/golden/make_spline_golden.py), plus reference-free
identities (product exactness, derivative vs scipy, integral vs quadrature)."""
import os

import numpy as np
import pytest
from scipy.interpolate import BSpline as SciBSpline

from omg_tools_b200.basics.spline import BSplineBasis, BSpline
from omg_tools_b200.basics import spline_extra as sx
from omg_tools_b200.basics.poly import new_symbol, Poly

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'spline_golden.npz'))
TOL = 1e-12


def basis(name):
    return BSplineBasis(G[name + '_knots'], int(G[name + '_degree']))


@pytest.mark.parametrize('name', ['b3', 'b2', 'b1', 'bz'])
def test_eval_derivative_greville(name):
    b = basis(name)
    assert np.abs(b(G[name + '_eval_x']) - G[name + '_eval']).max() < TOL
    assert np.allclose(b.greville(), G[name + '_greville'], atol=TOL)
    for o in range(1, b.degree + 1):
        Bd, P = b.derivative(o)
        assert np.array_equal(Bd.knots, G['%s_der%d_knots' % (name, o)])
        assert np.abs(P - G['%s_der%d_P' % (name, o)]).max() < 1e-9 * np.abs(P).max()


@pytest.mark.parametrize('name', ['b3', 'b2', 'b1'])
def test_shift_matrices(name):
    b = basis(name)
    assert np.abs(sx.shiftoverknot_T(b) - G[name + '_shiftoverknot_T']).max() < TOL
    assert np.abs(sx.extrapolate_T(b, 0.1) - G[name + '_extrapolate_T']).max() < TOL
    for ts in (0.0, 0.03, 0.0999):
        T, Tinv = sx.shiftfirstknot_T(b, ts, inverse=True)
        assert np.abs(T - G['%s_shiftfirst_T_%g' % (name, ts)]).max() < TOL
        assert np.abs(Tinv - G['%s_shiftfirst_Tinv_%g' % (name, ts)]).max() < 1e-9


def test_shiftoverknot_known_answer():
    # golden constant quoted from the reference
    T = sx.shiftoverknot_T(basis('b3'))
    assert np.allclose(T[0, :4], [0, .25, 7. / 12, 1. / 6])
    assert np.allclose(T[2, :4], [0, 0, 0, 1])
    assert np.allclose(T[12, 9:], [-1. / 6, 2 + 5. / 12, -9.25, 8])
    assert np.allclose(basis('b3').derivative(1)[1][0, :2], [-30, 30])


PAIRS = [('b1', 'b3'), ('b1', 'bz'), ('b1', 'b1'), ('b3', 'b3'), ('b2', 'b2'),
         ('b1', 'b2'), ('bz', 'bz')]


@pytest.mark.parametrize('n1,n2', PAIRS)
def test_sum_product(n1, n2):
    tag = '%s_%s' % (n1, n2)
    s1 = BSpline(basis(n1), G['c1_' + tag])
    s2 = BSpline(basis(n2), G['c2_' + tag])
    sm, pr = s1 + s2, s1 * s2
    assert np.array_equal(sm.basis.knots, G['sum_knots_' + tag])
    assert np.array_equal(pr.basis.knots, G['prod_knots_' + tag])
    assert np.abs(sm.coeffs - G['sum_coeffs_' + tag]).max() < TOL
    assert np.abs(pr.coeffs - G['prod_coeffs_' + tag]).max() < TOL
    # reference-free: product is exact pointwise
    x = np.linspace(0, 1, 97)
    assert np.abs(pr(x) - s1(x) * s2(x)).max() < 1e-12


def test_collision_row_expression():
    b1, b3 = basis('b1'), basis('b3')
    a0, a1, b = (BSpline(b1, G['row_' + k]) for k in ('ca0', 'ca1', 'cb'))
    x, y, e = (BSpline(b3, G['row_' + k]) for k in ('cx', 'cy', 'ce'))
    con = 0
    con += (a0 * 0.0 + a1 * 0.0) * 1.
    con += (a0 * x + a1 * y)
    con += (-b + 0.1 + 0.1 - e) * 1.
    assert con.basis.degree == int(G['row_degree'])
    assert np.array_equal(con.basis.knots, G['row_knots'])
    assert len(con.coeffs) == 41
    assert np.abs(con.coeffs - G['row_coeffs']).max() < 1e-11


def test_integrals_and_interval():
    b3 = basis('b3')
    s = BSpline(b3, G['int_c'])
    assert abs(s.integral() - float(G['int_value'])) < TOL
    ri = sx.running_integral(s)
    assert np.array_equal(ri.basis.knots, G['runint_knots'])
    assert np.abs(ri.coeffs - G['runint_coeffs']).max() < TOL
    Ti, ki = sx.get_interval_T(b3, 0.2, 0.7)
    assert np.abs(Ti - G['interval_T']).max() < TOL
    assert np.allclose(ki, G['interval_knots'])
    Tf, kn = sx.get_interval_T(BSplineBasis(G['trig_knots_theta'], 2), 0, 1.)
    assert np.abs(Tf - G['trig_Tf']).max() < TOL
    assert np.allclose(kn, G['trig_knots'])
    # reference-free: definite integral by quadrature
    xs = np.linspace(0.03, 1., 20001)
    quad = np.trapezoid(s(xs), xs)
    assert abs(sx.definite_integral(s, 0.03, 1.) - quad) < 1e-7


def test_derivative_vs_scipy():
    b3 = basis('b3')
    c = G['int_c']
    x = np.linspace(0.01, 0.99, 50)
    for o in (1, 2, 3):
        ours = BSpline(b3, c).derivative(o)(x)
        ref = SciBSpline(b3.knots, c, 3).derivative(o)(x)
        assert np.abs(ours - ref).max() < 1e-8 * max(1, np.abs(ref).max())


def test_evalspline_symbolic_matches_numeric():
    b3 = basis('b3')
    c = G['int_c']
    t = new_symbol('t', 'par')
    T = new_symbol('T', 'par')
    expr = sx.evalspline(BSpline(b3, c), t / T)
    assert isinstance(expr, Poly)
    for tv in (0., 0.31, 0.999):
        vals = {t.single_symbol(): tv, T.single_symbol(): 10.}
        num = sx.evalspline(BSpline(b3, c), tv / 10.)
        assert abs(expr.evaluate(vals) - num) < 1e-13
        assert abs(num - BSpline(b3, c)(tv / 10.)[0]) < 1e-13
