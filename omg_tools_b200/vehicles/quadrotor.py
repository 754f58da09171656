"""Planar quadrotor: the position (x, y) is the flat output, degree-4 splines;
thrust and pitch-rate limits become polynomial rows of degree 2 and 3 in the
spline coefficients (reference ``omgtools/vehicles/quadrotor.py``: bounds
32-36, trajectory constraints 48-62, initial/terminal constraints 64-84,
initial guess 97-105, parameters 116-122, signals 128-149, ode 154-157)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/quadrotor.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.shape import Circle
from ..basics.spline_extra import sample_splines


class Quadrotor(Vehicle):

    def __init__(self, radius=0.2, options=None, bounds=None):
        bounds = bounds or {}
        Vehicle.__init__(self, n_spl=2, degree=4, shapes=Circle(radius), options=options)
        self.radius = radius
        self.u1min = bounds.get('u1min', 2.)
        self.u1max = bounds.get('u1max', 15.)
        self.u2min = bounds.get('u2min', -8.)
        self.u2max = bounds.get('u2max', 8.)
        self.g = 9.81

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options['stop_tol'] = 1.e-2

    def define_trajectory_constraints(self, splines, horizon_time):
        T = horizon_time
        x, y = splines
        ddx, ddy = x.derivative(2), y.derivative(2)
        dddx, dddy = x.derivative(3), y.derivative(3)
        g_tf = self.g * (T**2)
        thrust2 = ddx**2 + (ddy + g_tf)**2
        self.define_constraint(-thrust2 + (T**4) * self.u1min**2, -inf, 0.)
        self.define_constraint(thrust2 - (T**4) * self.u1max**2, -inf, 0.)
        rate = dddx * (ddy + g_tf) - ddx * dddy
        self.define_constraint(-rate + thrust2 * (T * self.u2min), -inf, 0.)
        self.define_constraint(rate - thrust2 * (T * self.u2max), -inf, 0.)

    def get_initial_constraints(self, splines, horizon_time):
        T = horizon_time
        spl0 = self.define_parameter('spl0', 2)
        dspl0 = self.define_parameter('dspl0', 2)
        ddspl0 = self.define_parameter('ddspl0', 2)
        x, y = splines
        dx, dy = x.derivative(), y.derivative()
        ddx, ddy = x.derivative(2), y.derivative(2)
        return [(x, spl0[0]), (y, spl0[1]),
                (dx, T * dspl0[0]), (dy, T * dspl0[1]),
                (ddx, (T**2) * ddspl0[0]), (ddy, (T**2) * ddspl0[1])]

    def get_terminal_constraints(self, splines, horizon_time=None):
        position = self.define_parameter('poseT', 2)
        x, y = splines
        term_con = [(x, position[0]), (y, position[1])]
        term_con_der = []
        for d in range(1, self.degree + 1):
            term_con_der.extend([(x.derivative(d), 0.), (y.derivative(d), 0.)])
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        state = np.asarray(state, dtype=float)
        self.prediction['state'] = np.r_[state[:2], np.zeros(3)].T
        self.prediction['input'] = np.array([self.g, 0.])
        self.prediction['dspl'] = np.zeros(2)
        self.prediction['ddspl'] = np.zeros(2)

    def set_terminal_conditions(self, position):
        self.poseT = np.asarray(position, dtype=float)

    def get_init_spline_value(self, subgoals=None):
        L, d = len(self.basis), self.degree
        init_value = np.zeros((L, 2))
        pos0, posT = self.prediction['state'][:2], self.poseT
        for k in range(2):
            init_value[:, k] = np.r_[pos0[k] * np.ones(d),
                                     np.linspace(pos0[k], posT[k], L - 2 * d),
                                     posT[k] * np.ones(d)]
        return [init_value]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['state'][:2, -1] - self.poseT) > tol or
                np.linalg.norm(self.signals['dspl'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        parameters[self]['spl0'] = self.prediction['state'][:2]
        parameters[self]['dspl0'] = self.prediction['dspl']
        parameters[self]['ddspl0'] = self.prediction['ddspl']
        parameters[self]['poseT'] = self.poseT
        return parameters

    def define_collision_constraints(self, hyperplanes, room, splines, horizon_time):
        x, y = splines[0], splines[1]
        self.define_collision_constraints_2d(hyperplanes, room, [x, y], horizon_time)

    def splines2signals(self, splines, time):
        signals = {}
        x, y = splines[0], splines[1]
        dx, dy = x.derivative(), y.derivative()
        ddx, ddy = x.derivative(2), y.derivative(2)
        dddx, dddy = x.derivative(3), y.derivative(3)
        x_s, y_s = [np.asarray(v) for v in sample_splines([x, y], time)]
        dx_s, dy_s = [np.asarray(v) for v in sample_splines([dx, dy], time)]
        ddx_s, ddy_s = [np.asarray(v) for v in sample_splines([ddx, ddy], time)]
        dddx_s, dddy_s = [np.asarray(v) for v in sample_splines([dddx, dddy], time)]
        theta = np.arctan2(ddx_s, ddy_s + self.g)
        u1 = np.sqrt(ddx_s**2 + (ddy_s + self.g)**2)
        u2 = (dddx_s * (ddy_s + self.g) - ddx_s * dddy_s) / \
            ((ddy_s + self.g)**2 + ddx_s**2)
        signals['state'] = np.c_[x_s, y_s, dx_s, dy_s, theta].T
        signals['input'] = np.c_[u1, u2].T
        signals['dspl'] = np.c_[dx_s, dy_s].T
        signals['ddspl'] = np.c_[ddx_s, ddy_s].T
        return signals

    def state2pose(self, state):
        return np.r_[state[0], state[1], -state[4]]

    def ode(self, state, input):
        theta = state[4]
        u1, u2 = input[0], input[1]
        return np.r_[state[2:4], u1 * np.sin(theta), u1 * np.cos(theta) - self.g, u2].T
