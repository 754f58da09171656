"""Quadrotor3D: flat outputs f~ (thrust), q_phi = tan(phi/2), q_theta =
tan(theta/2) as degree-2 splines; accelerations tied to degree-4 slack
splines that are integrated twice to the position.

Model-building interface and row order of the reference's
``omgtools/vehicles/quadrotor3d.py`` (bounds 51-62, init 70-74, trajectory
constraints 76-133, initial/terminal constraints 135-174, initial guess
189-201, parameters 211-223, collision constraints 225-238, integrate_twice
240-254).  Rows are polynomials up to degree 5 in the decision variables.
"""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/quadrotor3d.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.poly import Poly, new_mid, collapse
from ..basics.shape import Sphere
from ..basics.spline import BSplineBasis, BSpline
from ..basics.spline_extra import evalspline, running_integral, sample_splines


class Quadrotor3D(Vehicle):

    def __init__(self, radius=0.2, options=None, bounds=None):
        bounds = bounds or {}
        Vehicle.__init__(self, n_spl=3, degree=2, shapes=Sphere(radius), options=options)
        self.u1min = bounds.get('u1min', 2.)
        self.u1max = bounds.get('u1max', 15.)
        self.u2min = bounds.get('u2min', -2.)
        self.u2max = bounds.get('u2max', 2.)
        self.u3min = bounds.get('u3min', -2.)
        self.u3max = bounds.get('u3max', 2.)
        self.phimin = bounds.get('phimin', -np.pi / 6)
        self.phimax = bounds.get('phimax', np.pi / 6)
        self.thetamin = bounds.get('thetamin', -np.pi / 6)
        self.thetamax = bounds.get('thetamax', np.pi / 6)
        self.g = 9.81
        self.radius = radius

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options['stop_tol'] = 5.e-1
        self.options['substitution'] = True
        self.options['exact_substitution'] = False

    def init(self):
        self.t = self.define_symbol('t')
        self.pos0 = self.define_parameter('pos0', 3)
        self.dpos0 = self.define_parameter('dpos0', 3)

    def define_trajectory_constraints(self, splines, horizon_time=None):
        # The position rows all depend on the 72 (94 for z) product-spline
        # coefficients of ddx/ddy/ddz through two running integrals re-anchored
        # at t/T.  CasADi keeps those coefficients as shared graph nodes; here
        # they are 'mid' symbols (basics/poly.py) and the rows stay affine in
        # them, lowering.py applies the chain rule.
        if horizon_time is None:
            horizon_time = self.define_symbol('T')
        T = horizon_time
        f_til, q_phi, q_theta = splines
        dq_phi, dq_theta = q_phi.derivative(), q_theta.derivative()
        # thrust, roll rate, pitch rate
        self.define_constraint(f_til * (1 + q_phi**2) * (1 + q_theta**2) - self.u1max, -inf, 0)
        self.define_constraint(-f_til * (1 + q_phi**2) * (1 + q_theta**2) + self.u1min, -inf, 0)
        self.define_constraint(2 * dq_phi - (1 + q_phi**2) * T * self.u2max, -inf, 0.)
        self.define_constraint(-2 * dq_phi + (1 + q_phi**2) * T * self.u2min, -inf, 0.)
        self.define_constraint(2 * dq_theta - (1 + q_theta**2) * T * self.u3max, -inf, 0.)
        self.define_constraint(-2 * dq_theta + (1 + q_theta**2) * T * self.u3min, -inf, 0.)
        # attitude limits
        self.define_constraint(q_phi - np.tan(0.5 * self.phimax), -inf, 0)
        self.define_constraint(-q_phi + np.tan(0.5 * self.phimin), -inf, 0)
        self.define_constraint(q_theta - np.tan(0.5 * self.thetamax), -inf, 0)
        self.define_constraint(-q_theta + np.tan(0.5 * self.thetamin), -inf, 0)
        if self.options['substitution']:
            ddx = f_til * (1 - q_phi**2) * (2 * q_theta)
            ddy = -f_til * (1 + q_theta**2) * (2 * q_phi)
            ddz = f_til * (1 - q_phi**2) * (1 - q_theta**2) - self.g
            if self.options['exact_substitution']:
                bx, by, bz = ddx.basis, ddy.basis, ddz.basis
            else:
                degree = 4
                knots = np.r_[np.zeros(degree), np.linspace(0., 1., 10 + 1), np.ones(degree)]
                bx = by = bz = BSplineBasis(knots, degree)
            self.ddx = self.define_spline_variable('ddx', 1, 1, basis=bx)[0]
            self.ddy = self.define_spline_variable('ddy', 1, 1, basis=by)[0]
            self.ddz = self.define_spline_variable('ddz', 1, 1, basis=bz)[0]
            # The reference does not shift these slacks over a knot (their names
            # lack 'seg', optilayer.py:470-490) and leaves the inconsistent warm
            # start to IPOPT's restoration phase.  This solver has no restoration
            # phase: shifting them with the position splines keeps the
            # warm start consistent; the NLP and its optimum are unchanged.
            for name in ('ddx', 'ddy', 'ddz'):
                self._splines_prim[name]['shift'] = True
            self.x, self.dx = self.integrate_twice(self.ddx, self.dpos0[0], self.pos0[0], self.t, T)
            self.y, self.dy = self.integrate_twice(self.ddy, self.dpos0[1], self.pos0[1], self.t, T)
            self.z, self.dz = self.integrate_twice(self.ddz, self.dpos0[2], self.pos0[2], self.t, T)
            if self.options['exact_substitution']:
                self.define_constraint(self.ddx - ddx, 0, 0)
                self.define_constraint(self.ddy - ddy, 0, 0)
                self.define_constraint(self.ddz - ddz, 0, 0)
            else:
                x, _ = self.integrate_twice(self._shared('ddx', ddx), self.dpos0[0],
                                            self.pos0[0], self.t, T)
                y, _ = self.integrate_twice(self._shared('ddy', ddy), self.dpos0[1],
                                            self.pos0[1], self.t, T)
                z, _ = self.integrate_twice(self._shared('ddz', ddz), self.dpos0[2],
                                            self.pos0[2], self.t, T)
                eps = 1e-3
                self.define_constraint(self.x - x, -eps, eps)
                self.define_constraint(self.y - y, -eps, eps)
                self.define_constraint(self.z - z, -eps, eps)

    def _shared(self, name, spline):
        """Spline whose nonlinear coefficients are named intermediates."""
        coeffs = np.empty(len(spline.coeffs), dtype=object)
        for k, c in enumerate(spline.coeffs):
            if isinstance(c, Poly) and c.degree() >= 2:
                c = new_mid('%s_%s_%d' % (self.label, name, k), c)
            coeffs[k] = c
        return BSpline(spline.basis, coeffs)

    def _positions(self, splines, horizon_time):
        f_til, q_phi, q_theta = splines
        if self.options['substitution']:
            return (self.x, self.y, self.z), (self.dx, self.dy, self.dz)
        ddx = f_til * (1 - q_phi**2) * (2 * q_theta)
        ddy = -f_til * (1 + q_theta**2) * (2 * q_phi)
        ddz = f_til * (1 - q_phi**2) * (1 - q_theta**2) - self.g
        x, dx = self.integrate_twice(ddx, self.dpos0[0], self.pos0[0], self.t, horizon_time)
        y, dy = self.integrate_twice(ddy, self.dpos0[1], self.pos0[1], self.t, horizon_time)
        z, dz = self.integrate_twice(ddz, self.dpos0[2], self.pos0[2], self.t, horizon_time)
        return (x, y, z), (dx, dy, dz)

    def get_initial_constraints(self, splines, horizon_time=None):
        f_til0 = self.define_parameter('f_til0', 1)
        q_phi0 = self.define_parameter('q_phi0', 1)
        q_theta0 = self.define_parameter('q_theta0', 1)
        self.define_parameter('dq_phi0', 1)
        self.define_parameter('dq_theta0', 1)
        f_til, q_phi, q_theta = splines
        return [(f_til, f_til0), (q_phi, q_phi0), (q_theta, q_theta0)]

    def get_terminal_constraints(self, splines, horizon_time=None):
        if horizon_time is None:
            horizon_time = self.define_symbol('T')
        posT = self.define_parameter('posT', 3)
        q_phiT = self.define_parameter('q_phiT', 1)
        q_thetaT = self.define_parameter('q_thetaT', 1)
        f_til, q_phi, q_theta = splines
        (x, y, z), (dx, dy, dz) = self._positions(splines, horizon_time)
        term_con = [(x, posT[0]), (y, posT[1]), (z, posT[2])]
        term_con_der = [(q_phi, q_phiT), (q_theta, q_thetaT), (f_til, self.g),
                        (dx, 0.), (dy, 0.), (dz, 0.)]
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.array([self.g, 0., 0.])
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)
        self.pose0 = np.r_[state[:3], state[6:], 0.]

    def set_terminal_conditions(self, position, roll=0, pitch=0):
        self.poseT = np.r_[position, roll, pitch, 0.].T

    def get_init_spline_value(self):
        L = len(self.basis)
        init_value = np.zeros((L, 3))
        q_phi0 = np.tan(self.prediction['state'][6] / 2.)
        q_theta0 = np.tan(self.prediction['state'][7] / 2.)
        init_value[:, 1] = np.linspace(q_phi0, np.tan(self.poseT[3] / 2.), L)
        init_value[:, 2] = np.linspace(q_theta0, np.tan(self.poseT[4] / 2.), L)
        return [init_value]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        return not (np.linalg.norm(self.signals['pose'][:3, -1] - self.poseT[:3]) > tol or
                    np.linalg.norm(self.signals['input'][:, -1]) - self.g > tol)

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        p = parameters[self]
        st, inp = self.prediction['state'], self.prediction['input']
        p['q_phi0'] = np.tan(st[6] / 2.)
        p['q_theta0'] = np.tan(st[7] / 2.)
        p['f_til0'] = inp[0] / ((1 + p['q_phi0']**2) * (1 + p['q_theta0']**2))
        p['dq_phi0'] = 0.5 * inp[1] * (1 + p['q_phi0']**2)
        p['dq_theta0'] = 0.5 * inp[2] * (1 + p['q_theta0']**2)
        p['pos0'] = st[:3]
        p['dpos0'] = st[3:6]
        p['posT'] = self.poseT[:3]
        p['q_phiT'] = np.tan(self.poseT[3] / 2.)
        p['q_thetaT'] = np.tan(self.poseT[4] / 2.)
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time=None):
        if horizon_time is None:
            horizon_time = self.define_symbol('T')
        (x, y, z), _ = self._positions(splines, horizon_time)
        self.define_collision_constraints_3d(hyperplanes, room, [x, y, z], horizon_time)

    def integrate_twice(self, ddx, dx0, x0, t, T=1.):
        """x(tau) with x(t/T) = x0, x'(t/T) = T dx0 ... as in the reference
        (quadrotor3d.py:240-254): two running integrals re-anchored at t/T."""
        symbolic = isinstance(t, Poly)
        ddx_int = T * running_integral(ddx)
        at = collapse(evalspline(ddx_int, t / T, True)) if symbolic else ddx_int(t / T)[0]
        dx = ddx_int - at + dx0
        dx_int = T * running_integral(dx)
        at = collapse(evalspline(dx_int, t / T, True)) if symbolic else dx_int(t / T)[0]
        x = dx_int - at + x0
        return x, dx

    def splines2signals(self, splines, time):
        """State (position, velocity, roll, pitch) and input (thrust, roll and
        pitch rate) trajectories of the flat outputs; the position is the
        double integral re-anchored at the predicted state at time[0]
        (reference quadrotor3d.py:253-275; the err_* plot signals are not
        produced)."""
        signals = {}
        f_til, q_phi, q_theta = splines[0], splines[1], splines[2]
        dq_phi, dq_theta = q_phi.derivative(), q_theta.derivative()
        ddx = f_til * (1 - q_phi**2) * (2 * q_theta)
        ddy = -f_til * (1 + q_theta**2) * (2 * q_phi)
        ddz = f_til * (1 - q_phi**2) * (1 - q_theta**2) - self.g
        st = self.prediction['state']
        x, dx = self.integrate_twice(ddx, st[3], st[0], time[0])
        y, dy = self.integrate_twice(ddy, st[4], st[1], time[0])
        z, dz = self.integrate_twice(ddz, st[5], st[2], time[0])
        x_s, y_s, z_s, dx_s, dy_s, dz_s = sample_splines([x, y, z, dx, dy, dz], time)
        f_til_s, q_phi_s, q_theta_s, dq_phi_s, dq_theta_s = sample_splines(
            [f_til, q_phi, q_theta, dq_phi, dq_theta], time)
        den = sample_splines([(1 + q_phi**2) * (1 + q_theta**2)], time)[0]
        phi = 2 * np.arctan2(q_phi_s, 1)
        theta = 2 * np.arctan2(q_theta_s, 1)
        dphi = 2 * np.array(dq_phi_s) / (1. + np.array(q_phi_s)**2)
        dtheta = 2 * np.array(dq_theta_s) / (1. + np.array(q_theta_s)**2)
        f_s = f_til_s * den
        signals['state'] = np.c_[x_s, y_s, z_s, dx_s, dy_s, dz_s, phi, theta].T
        signals['input'] = np.c_[f_s, dphi.T, dtheta.T].T
        return signals

    def state2pose(self, state):
        return np.r_[state[0], state[1], state[2], state[6], state[7], 0.]

    def ode(self, state, input):
        phi, theta = state[6], state[7]
        u1, u2, u3 = input[0], input[1], input[2]
        return np.r_[state[3:6], u1 * np.sin(theta) * np.cos(phi), -u1 * np.sin(phi),
                     -self.g + u1 * np.cos(phi) * np.cos(theta), u2, u3].T
