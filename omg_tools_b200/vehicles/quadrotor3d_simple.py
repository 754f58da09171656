"""3-D quadrotor, flat formulation: the position x, y, z as degree-4 splines; thrust, body
rates and tilt limits are quadratic rows in the second and third derivatives (reference
``omgtools/vehicles/quadrotor3d_simple.py``: bounds 30-47, trajectory constraints 56-77,
initial / terminal constraints 79-103, initial guess 114-122, parameters 133-146,
collision constraints 148-152, signals 154-181, ode 186-190).  No intermediates: every row
is a polynomial of degree <= 2 in the spline coefficients."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/quadrotor3d_simple.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.shape import Sphere
from ..basics.spline_extra import sample_splines


class SimpleQuadrotor3D(Vehicle):

    def __init__(self, radius=0.2, options=None, bounds=None):
        bounds = bounds or {}
        Vehicle.__init__(self, n_spl=3, degree=4, shapes=Sphere(radius), options=options)
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

    def init(self):
        self.T = self.define_symbol('T')

    def define_trajectory_constraints(self, splines, horizon_time):
        T = self.T
        x, y, z = splines
        ddx, ddy, ddz = x.derivative(2), y.derivative(2), z.derivative(2)
        dddx, dddy, dddz = x.derivative(3), y.derivative(3), z.derivative(3)
        az = ddz + self.g * (T**2)                 # vertical specific force, scaled by T^2
        thrust2 = ddx**2 + ddy**2 + az**2
        self.define_constraint(-thrust2 + (T**4) * self.u1min**2, -inf, 0.)
        self.define_constraint(thrust2 - (T**4) * self.u1max**2, -inf, 0.)
        self.define_constraint(-dddy * az + dddz * ddy - (az**2) * T * self.u2max, -inf, 0.)
        self.define_constraint(dddy * az - dddz * ddy + (az**2) * T * self.u2min, -inf, 0.)
        self.define_constraint(dddx * az - dddz * ddx - (az**2) * T * self.u3max, -inf, 0.)
        self.define_constraint(-dddx * az + dddz * ddx + (az**2) * T * self.u3min, -inf, 0.)
        self.define_constraint(-ddy - az * self.phimax, -inf, 0.)
        self.define_constraint(ddy + az * self.phimin, -inf, 0.)
        self.define_constraint(ddx - az * self.thetamax, -inf, 0.)
        self.define_constraint(-ddx + az * self.thetamin, -inf, 0.)

    def get_initial_constraints(self, splines, horizon_time):
        T = self.T
        spl0 = self.define_parameter('spl0', 3)
        dspl0 = self.define_parameter('dspl0', 3)
        ddspl0 = self.define_parameter('ddspl0', 3)
        x, y, z = splines
        dx, dy, dz = x.derivative(), y.derivative(), z.derivative()
        ddx, ddy, ddz = x.derivative(2), y.derivative(2), z.derivative(2)
        return [(x, spl0[0]), (y, spl0[1]), (z, spl0[2]),
                (dx, T * dspl0[0]), (dy, T * dspl0[1]), (dz, T * dspl0[2]),
                (ddx, (T**2) * ddspl0[0]), (ddy, (T**2) * ddspl0[1]), (ddz, (T**2) * ddspl0[2])]

    def get_terminal_constraints(self, splines, horizon_time=None):
        position = self.define_parameter('positionT', 3)
        x, y, z = splines
        term_con = [(x, position[0]), (y, position[1]), (z, position[2])]
        term_con_der = []
        for d in range(1, self.degree + 1):
            term_con_der.extend([(x.derivative(d), 0.), (y.derivative(d), 0.),
                                 (z.derivative(d), 0.)])
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.array([self.g, 0., 0.])
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)

    def set_terminal_conditions(self, position):
        self.positionT = np.asarray(position, dtype=float)

    def get_init_spline_value(self, subgoals=None):
        L = len(self.basis)
        init_value = np.zeros((L, 3))
        for k in range(3):
            init_value[:, k] = np.linspace(self.prediction['state'][k], self.positionT[k], L)
        return [init_value]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['pose'][:3, -1] - self.positionT) > tol or
                np.linalg.norm(self.signals['dspl'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        p, st = parameters[self], self.prediction['state']
        f0, phi0, theta0 = self.prediction['input'][0], st[6], st[7]
        p['spl0'] = st[:3]
        p['ddspl0'] = [f0 * np.cos(phi0) * np.sin(theta0), -f0 * np.sin(phi0),
                       f0 * np.cos(phi0) * np.cos(theta0) - self.g]
        p['dspl0'] = st[3:6]
        p['positionT'] = self.positionT
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time):
        x, y, z = splines[0], splines[1], splines[2]
        self.define_collision_constraints_3d(hyperplanes, room, [x, y, z], horizon_time)

    def splines2signals(self, splines, time):
        signals = {}
        x, y, z = splines[0], splines[1], splines[2]
        der = lambda d: [np.asarray(v) for v in sample_splines(
            [x.derivative(d), y.derivative(d), z.derivative(d)], time)]
        x_s, y_s, z_s = [np.asarray(v) for v in sample_splines([x, y, z], time)]
        dx_s, dy_s, dz_s = der(1)
        ddx_s, ddy_s, ddz_s = der(2)
        dddx_s, dddy_s, dddz_s = der(3)
        az = ddz_s + self.g
        phi = np.arctan2(-ddy_s, np.sqrt(ddx_s**2 + az**2))
        theta = np.arctan2(ddx_s, az)
        u1 = np.sqrt(ddx_s**2 + ddy_s**2 + az**2)
        u2 = (-dddy_s * (ddx_s**2 + az**2) + ddy_s * (ddx_s * dddx_s + dddz_s * az)) / \
            ((ddx_s**2 + ddy_s**2 + az**2) * np.sqrt(ddx_s**2 + az**2))
        u3 = (az * dddx_s - ddx_s * dddz_s) / (az**2 + ddx_s**2)
        signals['state'] = np.c_[x_s, y_s, z_s, dx_s, dy_s, dz_s, phi, theta].T
        signals['input'] = np.c_[u1, u2, u3].T
        signals['dspl'] = np.c_[dx_s, dy_s, dz_s].T
        signals['ddspl'] = np.c_[ddx_s, ddy_s, ddz_s].T
        signals['dddspl'] = np.c_[dddx_s, dddy_s, dddz_s].T
        return signals

    def state2pose(self, state):
        return np.r_[state[0], state[1], state[2], state[6], state[7], 0.]

    def ode(self, state, input):
        phi, theta = state[6], state[7]
        u1, u2, u3 = input[0], input[1], input[2]
        return np.r_[state[3:6], u1 * np.sin(theta) * np.cos(phi), -u1 * np.sin(phi),
                     -self.g + u1 * np.cos(phi) * np.cos(theta), u2, u3].T
