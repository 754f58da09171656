"""Dubins vehicle: forward speed and heading, flat outputs v~ = v / (1 + tan^2(theta/2))
and tan(theta/2) as degree-3 splines (reference ``omgtools/vehicles/dubins.py``: bounds
38-45, trajectory constraints 72-126, initial / terminal constraints 155-193, initial
guess 206-214, parameters 225-233, collision constraints 235-251, integrate_once
253-259, signals 261-288).

The position is the running integral of v~(1 - tg^2), 2 v~ tg re-anchored at t/T; the
coefficients of these product splines are shared intermediates (basics/poly.py).  All three
formulations of the reference are built: the default (no substitution: the integrated
position enters the terminal and collision rows directly, so the hyperplane normal -- and
(1 + tg^2) for a non-circular shape -- multiplies the intermediates: rows affine in them
with x-dependent coefficients, lowering.py cross-Hessian slots), ``substitution`` (slack
velocity splines dx, dy carry the position and the band |int(dx) - int(v~(1-tg^2))| <= 1e-3
ties them to the flat outputs; examples/p2p_dubins.py) and ``exact_substitution`` (dx, dy on
the product basis, tied by equality rows)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/dubins.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.poly import Poly, new_mid, collapse, rel_time
from ..basics.shape import Circle
from ..basics.spline import BSplineBasis, BSpline
from ..basics.spline_extra import evalspline, running_integral, sample_splines


class Dubins(Vehicle):

    def __init__(self, shapes=None, options=None, bounds=None):
        bounds = bounds or {}
        shapes = shapes if shapes is not None else Circle(0.1)
        degree = options['degree'] if options is not None and 'degree' in options else 3
        Vehicle.__init__(self, n_spl=2, degree=degree, shapes=shapes, options=options)
        self.vmax = bounds.get('vmax', 0.5)
        self.amax = bounds.get('amax', 1.)
        self.wmin = bounds.get('wmin', -np.pi / 6.)
        self.wmax = bounds.get('wmax', np.pi / 6.)

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options.update({'stop_tol': 1.e-2, 'substitution': False,
                             'exact_substitution': False, 'init_v_til': 0.})

    def init(self):
        self.t = self.define_symbol('t')
        self.pos0 = self.define_parameter('pos0', 2)

    def _shared(self, name, spline):
        coeffs = np.empty(len(spline.coeffs), dtype=object)
        for k, c in enumerate(spline.coeffs):
            if isinstance(c, Poly) and c.degree() >= 2:
                c = new_mid('%s_%s_%d' % (self.label, name, k), c)
            coeffs[k] = c
        return BSpline(spline.basis, coeffs)

    def _flat_position(self, splines, horizon_time):
        """x, y as running integrals of the flat-output products; the product-spline
        coefficients are shared intermediates, created once per problem construction."""
        key = (id(splines[0].coeffs), id(splines[1].coeffs))
        cache = self.__dict__.setdefault('_flat_cache', {})
        if key not in cache:
            v_til, tg_ha = splines
            dx = v_til * (1 - tg_ha**2)
            dy = v_til * (2 * tg_ha)
            tag = '' if not cache else str(len(cache))      # a trailer's copies: own mids
            if not self.options['substitution']:
                # the coefficients of the integrated, re-anchored POSITION splines are the
                # shared intermediates: the rows that use the position (terminal rows,
                # hyperplane normal x position, (1 + tg^2) x position for a non-circular
                # shape) then hold one intermediate per monomial instead of the sum over
                # every product-spline coefficient before that knot -- 10-20 x smaller tables
                x = self._shared('x' + tag, self.integrate_once(dx, self.pos0[0], self.t, horizon_time))
                y = self._shared('y' + tag, self.integrate_once(dy, self.pos0[1], self.t, horizon_time))
            else:       # substitution: the position enters two band rows only -- the
                        # product-spline coefficients are the smaller set
                x = self.integrate_once(self._shared('dx' + tag, dx), self.pos0[0], self.t, horizon_time)
                y = self.integrate_once(self._shared('dy' + tag, dy), self.pos0[1], self.t, horizon_time)
            cache[key] = (splines, x, y)   # (the splines keep the ids alive)
        return cache[key][1], cache[key][2]

    def define_trajectory_constraints(self, splines, horizon_time):
        T = horizon_time
        v_til, tg_ha = splines
        dtg_ha = tg_ha.derivative()
        self.define_constraint(v_til * (1 + tg_ha**2) - self.vmax, -inf, 0.)
        self.define_constraint(-v_til, -inf, 0)          # forward driving only
        if self.options['substitution']:
            dx = v_til * (1 - tg_ha**2)
            dy = v_til * (2 * tg_ha)
            if self.options['exact_substitution']:
                self.dx = self.define_spline_variable('dx', 1, 1, basis=dx.basis)[0]
                self.dy = self.define_spline_variable('dy', 1, 1, basis=dy.basis)[0]
                self.x = self.integrate_once(self.dx, self.pos0[0], self.t, T)
                self.y = self.integrate_once(self.dy, self.pos0[1], self.t, T)
                self.define_constraint(self.dx - dx, 0, 0)
                self.define_constraint(self.dy - dy, 0, 0)
            else:
                degree = 3
                knots = np.r_[np.zeros(degree), np.linspace(0., 1., 10 + 1), np.ones(degree)]
                basis = BSplineBasis(knots, degree)
                self.dx = self.define_spline_variable('dx', 1, 1, basis=basis)[0]
                self.dy = self.define_spline_variable('dy', 1, 1, basis=basis)[0]
                for name in ('dx', 'dy'):   # keep the warm start consistent (see quadrotor3d.py)
                    self._splines_prim[name]['shift'] = True
                self.x = self.integrate_once(self.dx, self.pos0[0], self.t, T)
                self.y = self.integrate_once(self.dy, self.pos0[1], self.t, T)
                x, y = self._flat_position(splines, T)
                eps = 1e-3
                self.define_constraint(self.x - x, -eps, eps)
                self.define_constraint(self.y - y, -eps, eps)
        self.define_constraint(2 * dtg_ha - (1 + tg_ha**2) * T * self.wmax, -inf, 0.)
        self.define_constraint(-2 * dtg_ha + (1 + tg_ha**2) * T * self.wmin, -inf, 0.)

    def get_initial_constraints(self, splines, horizon_time):
        v_til0 = self.define_parameter('v_til0', 1)
        tg_ha0 = self.define_parameter('tg_ha0', 1)
        dtg_ha0 = self.define_parameter('dtg_ha0', 1)
        v_til, tg_ha = splines
        return [(v_til, v_til0), (tg_ha, tg_ha0),
                (tg_ha.derivative(), horizon_time * dtg_ha0)]

    def get_terminal_constraints(self, splines, horizon_time=None):
        posT = self.define_parameter('posT', 2)
        tg_haT = self.define_parameter('tg_haT', 1)
        v_til, tg_ha = splines
        if self.options['substitution']:
            x, y = self.x, self.y
        else:
            if horizon_time is None:
                horizon_time = self.define_symbol('T')
            x, y = self._flat_position(splines, horizon_time)
        term_con = [(x, posT[0]), (y, posT[1]), (tg_ha, tg_haT)]
        term_con_der = [(v_til, 0.), (tg_ha.derivative(), 0.)]
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.zeros(2)
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)
        self.pose0 = np.asarray(state, dtype=float)

    def set_terminal_conditions(self, pose):
        self.poseT = np.asarray(pose, dtype=float)

    def get_init_spline_value(self, subgoals=None):
        L = len(self.basis)
        init_value = np.zeros((L, 2))
        # The reference starts from v~ = 0 (dubins.py:209-214), a point where the position
        # does not depend on the heading; IPOPT leaves it through its restoration phase,
        # this solver does not.  Option 'init_v_til' (default 0 = the
        # reference's guess) sets a rolling initial guess instead.
        init_value[:, 0] = self.options.get('init_v_til', 0.)
        tg_ha0 = np.tan(self.prediction['state'][2] / 2.)
        tg_haT = np.tan(self.poseT[2] / 2.)
        init_value[:, 1] = np.linspace(tg_ha0, tg_haT, L)
        return [init_value]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['state'][:, -1] - self.poseT) > tol or
                np.linalg.norm(self.signals['input'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        p = parameters[self]
        p['tg_ha0'] = np.tan(self.prediction['state'][2] / 2.)
        p['v_til0'] = self.prediction['input'][0] / (1 + p['tg_ha0']**2)
        p['dtg_ha0'] = 0.5 * self.prediction['input'][1] * (1 + p['tg_ha0']**2)
        p['pos0'] = self.prediction['state'][:2]
        p['posT'] = self.poseT[:2]
        p['tg_haT'] = np.tan(self.poseT[2] / 2.)
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time):
        if self.options['substitution']:
            x, y = self.x, self.y
        else:
            x, y = self._flat_position(splines, horizon_time)
        if isinstance(self.shapes[0], Circle):     # heading irrelevant for a disc
            self.define_collision_constraints_2d(hyperplanes, room, [x, y], horizon_time)
        else:
            self.define_collision_constraints_2d(hyperplanes, room, [x, y], horizon_time,
                                                 tg_ha=splines[1])

    def integrate_once(self, dx, x0, t, T=1.):
        """x(tau) with x(t/T) = x0 (reference dubins.py:253-259)."""
        dx_int = T * running_integral(dx)
        if isinstance(t, Poly):
            return dx_int - collapse(evalspline(dx_int, rel_time(t, T), True)) + x0
        return dx_int - dx_int(t / T)[0] + x0

    def splines2signals(self, splines, time):
        signals = {}
        v_til, tg_ha = splines[0], splines[1]
        dtg_ha = tg_ha.derivative()
        dx = v_til * (1 - tg_ha**2)
        dy = v_til * (2 * tg_ha)
        st = self.prediction['state']
        x = self.integrate_once(dx, st[0], time[0])
        y = self.integrate_once(dy, st[1], time[0])
        x_s, y_s, v_til_s, tg_ha_s, dtg_ha_s = [
            np.asarray(v) for v in sample_splines([x, y, v_til, tg_ha, dtg_ha], time)]
        den = np.asarray(sample_splines([(1 + tg_ha**2)], time)[0])
        theta = 2 * np.arctan2(tg_ha_s, 1)
        dtheta = 2 * dtg_ha_s / (1. + tg_ha_s**2)
        signals['state'] = np.c_[x_s, y_s, theta].T
        signals['input'] = np.c_[v_til_s * den, dtheta].T
        acc = (v_til * (1 + tg_ha**2)).derivative()
        signals['acc'] = np.c_[np.asarray(sample_splines([acc], time)[0])].T
        return signals

    def state2pose(self, state):
        return state

    def ode(self, state, input):
        theta, v, w = state[2], input[0], input[1]
        return np.r_[v * np.cos(theta), v * np.sin(theta), w].T
