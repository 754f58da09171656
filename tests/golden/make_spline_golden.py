"""Generate tests/golden/spline_golden.npz from the REFERENCE's own spline code.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_spline_golden.py

The reference's ``omgtools/basics/spline.py`` and ``spline_extra.py`` are pure
numpy/scipy except for their ``import casadi``; CasADi is not installed, so a
dummy ``casadi`` module exposing just the imported names is registered and the
two files are loaded straight from /root/reference (nothing is copied).  The
matrices written here pin this framework's spline algebra (tests/test_spline.py)
to the reference's numerics.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = '/root/reference/omgtools/basics'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                   'spline_golden.npz')


def load_reference():
    cas = types.ModuleType('casadi')
    for name in ('MX', 'SX', 'DM'):
        setattr(cas, name, type(name, (), {}))
    for name in ('mtimes', 'Function', 'vertcat'):
        setattr(cas, name, lambda *a, **k: None)
    cas.inf = np.inf
    sys.modules['casadi'] = cas
    pkg = types.ModuleType('refbasics')
    pkg.__path__ = [REF]
    sys.modules['refbasics'] = pkg
    mods = {}
    for name in ('spline', 'spline_extra'):
        spec = importlib.util.spec_from_file_location(
            'refbasics.' + name, os.path.join(REF, name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules['refbasics.' + name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods['spline'], mods['spline_extra']


def main():
    rs, rx = load_reference()
    out = {}
    rng = np.random.default_rng(0)

    def veh_knots(deg, n_int):
        return np.r_[np.zeros(deg), np.linspace(0., 1., n_int + 1), np.ones(deg)]

    b3 = rs.BSplineBasis(veh_knots(3, 10), 3)     # holonomic vehicle basis
    b2 = rs.BSplineBasis(veh_knots(2, 10), 2)     # quadrotor3d basis
    b1 = rs.BSplineBasis(veh_knots(1, 10), 1)     # hyperplane basis
    bz = rs.BSplineBasis([0, 0, 0, 1, 1, 1], 2)   # obstacle position basis
    bases = {'b3': b3, 'b2': b2, 'b1': b1, 'bz': bz}
    for name, b in bases.items():
        out[name + '_knots'] = b.knots
        out[name + '_degree'] = b.degree
        x = np.linspace(0, 1, 37)
        out[name + '_eval_x'] = x
        out[name + '_eval'] = b(x).toarray()
        out[name + '_greville'] = np.array(b.greville())
        for o in range(1, b.degree + 1):
            Bd, P = b.derivative(o)
            out['%s_der%d_knots' % (name, o)] = Bd.knots
            out['%s_der%d_P' % (name, o)] = P.toarray()
        if name == 'bz':      # no interior knots: horizon shift undefined
            continue
        out[name + '_shiftoverknot_T'] = rx.shiftoverknot_T(b)
        out[name + '_extrapolate_T'] = rx.extrapolate_T(b, 0.1)
        for ts in (0.0, 0.03, 0.0999):
            T, Tinv = rx.shiftfirstknot_T(b, ts, inverse=True)
            out['%s_shiftfirst_T_%g' % (name, ts)] = T
            out['%s_shiftfirst_Tinv_%g' % (name, ts)] = Tinv

    # sums and products (basis knots, result coefficients on random inputs)
    pairs = [('b1', 'b3'), ('b1', 'bz'), ('b1', 'b1'), ('b3', 'b3'),
             ('b2', 'b2'), ('b1', 'b2'), ('bz', 'bz')]
    for n1, n2 in pairs:
        B1, B2 = bases[n1], bases[n2]
        c1, c2 = rng.standard_normal(len(B1)), rng.standard_normal(len(B2))
        s1, s2 = rs.BSpline(B1, c1), rs.BSpline(B2, c2)
        tag = '%s_%s' % (n1, n2)
        out['c1_' + tag], out['c2_' + tag] = c1, c2
        sm, pr = s1 + s2, s1 * s2
        out['sum_knots_' + tag], out['sum_coeffs_' + tag] = sm.basis.knots, sm.coeffs
        out['prod_knots_' + tag], out['prod_coeffs_' + tag] = pr.basis.knots, pr.coeffs
    # nested expression with the shape of the vehicle-side collision row
    # a0*x + a1*y - b + r - eps  (reference vehicle.py:147-158)
    ca0, ca1, cb = (rng.standard_normal(len(b1)) for _ in range(3))
    cx, cy, ce = (rng.standard_normal(len(b3)) for _ in range(3))
    con = 0
    con += (rs.BSpline(b1, ca0) * 0.0 + rs.BSpline(b1, ca1) * 0.0) * 1.
    con += (rs.BSpline(b1, ca0) * rs.BSpline(b3, cx) +
            rs.BSpline(b1, ca1) * rs.BSpline(b3, cy))
    con += (-rs.BSpline(b1, cb) + 0.1 + 0.1 - rs.BSpline(b3, ce)) * 1.
    for k, v in dict(ca0=ca0, ca1=ca1, cb=cb, cx=cx, cy=cy, ce=ce).items():
        out['row_' + k] = v
    out['row_knots'], out['row_coeffs'] = con.basis.knots, con.coeffs
    out['row_degree'] = con.basis.degree

    # integral, running integral, crop / interval, knot insertion
    c = rng.standard_normal(len(b3))
    s = rs.BSpline(b3, c)
    out['int_c'] = c
    out['int_value'] = s.integral()
    ri = rx.running_integral(s)
    out['runint_knots'], out['runint_coeffs'] = ri.basis.knots, ri.coeffs
    Ti, ki = rx.get_interval_T(b3, 0.2, 0.7)
    out['interval_T'], out['interval_knots'] = Ti, np.array(ki)
    # rotating-obstacle trig basis (reference obstacle.py:309-314), omega=1.5*2pi/10
    Ts, T = 10. / 1.5, 10.
    nq = int(np.ceil(4 * T / Ts))
    knots_theta = np.r_[np.zeros(3), np.hstack(
        [0.25 * k * np.ones(2) for k in range(1, nq + 1)]), 0.25 * nq] * (Ts / T)
    Tf, kn = rx.get_interval_T(rs.BSplineBasis(knots_theta, 2), 0, 1.)
    out['trig_knots_theta'], out['trig_Tf'], out['trig_knots'] = knots_theta, Tf, np.array(kn)
    np.savez_compressed(OUT, **out)
    print('wrote %s (%d arrays)' % (OUT, len(out)))


if __name__ == '__main__':
    main()
