"""Bicycle (kinematic car with a steering angle): flat outputs v~ = V / (1 + tan^2(theta/2))
and tan(theta/2) as degree-2 splines, steering angle from tan(delta) = 2 L d(tg) /
(v~ (1 + tg^2)^2) (reference ``omgtools/vehicles/bicycle.py``: bounds 55-66, trajectory
constraints 78-133, l'Hopital start constraint 135-165, terminal constraints 167-192,
parameters 226-257, collision constraints 259-278, signals 287-357, ode 362-369).

Lowering: the steering-rate rows contain v~^2 (1 + tg^2)^4, a polynomial of degree 10 in
the spline coefficients.  CasADi keeps its factors as shared graph nodes; here the
product splines q = v~ (1 + tg^2)^2, r = (1 + tg^2)^2 and s = v~ tg (1 + tg^2) are shared
intermediates (basics/poly.py), so that the rows read  2 L ddtg q - 2 L dtg (dv r + 4 s dtg)
- (T^2 q^2 + (2 L dtg)^2) ddelta_max: monomials with at most two intermediates (lowering.py:
mid-mid Hessian slots).  The position is the integrated product spline as for Dubins."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/bicycle.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .dubins import Dubins
from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.poly import rel_time
from ..basics.shape import Circle
from ..basics.spline import BSplineBasis
from ..basics.spline_extra import evalspline, sample_splines


class Bicycle(Dubins):

    def __init__(self, length=0.4, options=None, bounds=None, shapes=None):
        bounds = bounds or {}
        Vehicle.__init__(self, n_spl=2, degree=2,
                         shapes=shapes if shapes is not None else Circle(length / 2.),
                         options=options)
        self.vmax = bounds.get('vmax', 0.8)
        self.amax = bounds.get('amax', 1.)
        self.dmin = bounds.get('dmin', -np.pi / 6.)       # steering angle [rad]
        self.dmax = bounds.get('dmax', np.pi / 6.)
        self.ddmin = bounds.get('ddmin', -np.pi / 4.)     # steering rate [rad/s]
        self.ddmax = bounds.get('ddmax', np.pi / 4.)
        self.length = length
        self.steer_sign = 1.                              # AGV: rear-wheel steering, -1

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options.update({'plot_type': 'bicycle', 'substitution': False,
                             'exact_substitution': False, 'init_v_til': 0.})

    def init(self):
        self.t = self.define_symbol('t')
        self.pos0 = self.define_symbol('pos0', 2)      # resolved to the parameter below

    def define_trajectory_constraints(self, splines, horizon_time):
        T, L, sg = horizon_time, self.length, self.steer_sign
        v_til, tg_ha = splines
        dv_til, dtg_ha = v_til.derivative(), tg_ha.derivative()
        ddtg_ha = tg_ha.derivative(2)
        self.define_constraint(v_til * (1 + tg_ha**2) - self.vmax, -inf, 0.)
        self.define_constraint(
            dv_til * (1 + tg_ha**2) + 2 * v_til * tg_ha * dtg_ha - T * self.amax, -inf, 0.)
        if self.options['substitution']:
            dx = v_til * (1 - tg_ha**2)
            dy = v_til * (2 * tg_ha)
            if self.options['exact_substitution']:
                self.dx = self.define_spline_variable('dx', 1, 1, basis=dx.basis)[0]
                self.dy = self.define_spline_variable('dy', 1, 1, basis=dy.basis)[0]
                self.define_constraint(self.dx - dx, 0., 0.)
                self.define_constraint(self.dy - dy, 0., 0.)
                self.x = self.integrate_once(self.dx, self.pos0[0], self.t, T)
                self.y = self.integrate_once(self.dy, self.pos0[1], self.t, T)
            else:
                degree = 2
                knots = np.r_[np.zeros(degree), np.linspace(0., 1., 10 + 1), np.ones(degree)]
                basis = BSplineBasis(knots, degree)
                self.dx = self.define_spline_variable('dx', 1, 1, basis=basis)[0]
                self.dy = self.define_spline_variable('dy', 1, 1, basis=basis)[0]
                for name in ('dx', 'dy'):
                    self._splines_prim[name]['shift'] = True
                self.x = self.integrate_once(self.dx, self.pos0[0], self.t, T)
                self.y = self.integrate_once(self.dy, self.pos0[1], self.t, T)
                x, y = self._flat_position(splines, T)
                eps = 1e-2
                self.define_constraint(self.x - x, -eps, eps)
                self.define_constraint(self.y - y, -eps, eps)
        # shared product splines (see module docstring)
        q = self._shared('q', v_til * (1 + tg_ha**2)**2)
        r = self._shared('r', (1 + tg_ha**2)**2)
        s = self._shared('s', v_til * tg_ha * (1 + tg_ha**2))
        # steering angle: tan(delta) = sg 2 L dtg / q within [tan(dmin), tan(dmax)]
        self.define_constraint(sg * 2 * dtg_ha * L - q * np.tan(self.dmax) * T, -inf, 0.)
        self.define_constraint(-sg * 2 * dtg_ha * L + q * np.tan(self.dmin) * T, -inf, 0.)
        # steering rate: d(delta)/dt = num / den within [ddmin, ddmax]
        num = 2 * L * ddtg_ha * q - 2 * L * dtg_ha * (dv_til * r + 4 * s * dtg_ha)
        den = (T**2) * q * q + (2 * L * dtg_ha)**2
        self.define_constraint(sg * num - den * self.ddmax, -inf, 0.)
        self.define_constraint(-sg * num + den * self.ddmin, -inf, 0.)
        self.define_constraint(-v_til, -inf, 0)           # forward driving only

    def get_initial_constraints(self, splines, horizon_time):
        T, L, sg = horizon_time, self.length, self.steer_sign
        v_til0 = self.define_parameter('v_til0', 1)
        tg_ha0 = self.define_parameter('tg_ha0', 1)
        dtg_ha0 = self.define_parameter('dtg_ha0', 1)
        v_til, tg_ha = splines
        dv_til, dtg_ha = v_til.derivative(), tg_ha.derivative()
        ddtg_ha = tg_ha.derivative(2)
        hop0 = self.define_parameter('hop0', 1)
        tdelta0 = self.define_parameter('tdelta0', 1)     # tan(delta0)
        # l'Hopital on tan(delta) = 2 L dtg / (v~ (1+tg^2)^2) when starting from standstill
        t0 = rel_time(self.t, T)
        self.define_constraint(
            hop0 * (sg * 2. * evalspline(ddtg_ha, t0) * L
                    - tdelta0 * (evalspline(dv_til, t0) * (1. + tg_ha0**2)**2) * T), 0., 0.)
        return [(v_til, v_til0), (tg_ha, tg_ha0), (dtg_ha, T * dtg_ha0)]

    def get_terminal_constraints(self, splines, horizon_time=None):
        T = horizon_time if horizon_time is not None else self.define_symbol('T')
        posT = self.define_parameter('posT', 2)
        v_tilT = self.define_parameter('v_tilT', 1)
        dv_tilT = self.define_parameter('dv_tilT', 1)
        tg_haT = self.define_parameter('tg_haT', 1)
        dtg_haT = self.define_parameter('dtg_haT', 1)
        ddtg_haT = self.define_parameter('ddtg_haT', 1)
        self.define_parameter('pos0', 2)                  # anchor of the integration
        v_til, tg_ha = splines
        dv_til, dtg_ha = v_til.derivative(), tg_ha.derivative()
        ddtg_ha = tg_ha.derivative(2)
        if self.options['substitution']:
            x, y = self.x, self.y
        else:
            x, y = self._flat_position(splines, T)
        term_con = [(x, posT[0]), (y, posT[1]), (tg_ha, tg_haT)]
        term_con_der = [(v_til, v_tilT), (dtg_ha, T * dtg_haT),
                        (dv_til, dv_tilT), (ddtg_ha, T**2 * ddtg_haT)]
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.zeros(2)
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)
        self.pose0 = self.prediction['state'][:3]
        self.delta0 = self.prediction['state'][3]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['state'][:3, -1] - self.poseT) > tol or
                np.linalg.norm(self.signals['input'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        p, st = parameters[self], self.prediction['state']
        p['tg_ha0'] = np.tan(st[2] / 2)
        p['v_til0'] = self.prediction['input'][0] / (1 + p['tg_ha0']**2)
        p['pos0'] = st[:2]
        p['posT'] = self.poseT[:2]
        p['v_tilT'] = 0.
        p['dv_tilT'] = 0.
        p['tg_haT'] = np.tan(self.poseT[2] / 2)
        p['dtg_haT'] = 0.
        p['ddtg_haT'] = 0.
        if p['v_til0'] <= 1e-4:          # standstill: impose the steering angle by l'Hopital
            p['hop0'] = 1.
            p['v_til0'] = 0.
            p['dtg_ha0'] = 0.
            p['tdelta0'] = np.tan(st[3])
        else:
            p['hop0'] = 0.
            p['dtg_ha0'] = self.steer_sign * np.tan(st[3]) * p['v_til0'] * \
                (1 + p['tg_ha0']**2)**2 / (2 * self.length)
        return parameters

    def splines2signals(self, splines, time):
        signals = {}
        L, sg = self.length, self.steer_sign
        v_til, tg_ha = splines[0], splines[1]
        dv_til, dtg_ha = v_til.derivative(), tg_ha.derivative()
        ddtg_ha = tg_ha.derivative(2)
        dx = v_til * (1 - tg_ha**2)
        dy = v_til * (2 * tg_ha)
        st = self.prediction['state']
        x = self.integrate_once(dx, st[0], time[0])
        y = self.integrate_once(dy, st[1], time[0])
        tg_ha, v_til, dtg_ha, dv_til, ddtg_ha = [
            np.array(sample_splines([spl], time)) for spl in (tg_ha, v_til, dtg_ha, dv_til, ddtg_ha)]
        theta = 2 * np.arctan2(tg_ha, 1)
        delta = np.arctan2(sg * 2 * dtg_ha * L, v_til * (1 + tg_ha**2)**2)
        with np.errstate(divide='ignore', invalid='ignore'):
            ddelta = sg * (2 * ddtg_ha * L * (v_til * (1 + tg_ha**2)**2)
                           - 2 * dtg_ha * L * (dv_til * (1 + tg_ha**2)**2
                                               + v_til * (4 * tg_ha + 4 * tg_ha**3) * dtg_ha)) / \
                (v_til**2 * (1 + tg_ha**2)**4 + (2 * dtg_ha * L)**2)
        # where v~ and dtg vanish the quotient is 0/0: l'Hopital / neighbouring samples,
        # exactly as the reference does (bicycle.py:312-330)
        if v_til[0, 0] <= 1e-4 and dtg_ha[0, 0] <= 1e-4:
            delta[0, 0] = np.arctan2(sg * 2 * ddtg_ha[0, 0] * L, dv_til[0, 0] * (1 + tg_ha[0, 0]**2)**2)
            ddelta[0, 0] = ddelta[0, 1]
        for k in range(1, len(time) - 1):
            if v_til[0, k] <= 1e-3 and dtg_ha[0, k] <= 1e-3:
                if ddtg_ha[0, k] <= 1e-4 and dv_til[0, k] <= 1e-4:
                    delta[0, k] = delta[0, k - 1]
                else:
                    delta[0, k] = np.arctan2(sg * 2 * ddtg_ha[0, k] * L,
                                             dv_til[0, k] * (1 + tg_ha[0, k]**2)**2)
                ddelta[0, k] = ddelta[0, k - 1]
        if v_til[0, -1] <= 1e-4 and dtg_ha[0, -1] <= 1e-4:
            delta[0, -1] = delta[0, -2]
            ddelta[0, -1] = ddelta[0, -2]
        x_s, y_s = sample_splines([x, y], time)
        signals['state'] = np.r_[np.c_[x_s, y_s].T, theta, delta]
        signals['input'] = np.r_[np.c_[v_til * (1 + tg_ha**2)], ddelta]
        signals['delta'] = delta
        return signals

    def state2pose(self, state):
        return state[:3]

    def ode(self, state, input):
        u1, u2 = input[0], input[1]
        return np.r_[u1 * np.cos(state[2]), u1 * np.sin(state[2]),
                     self.steer_sign * u1 / self.length * np.tan(state[3]), u2].T
