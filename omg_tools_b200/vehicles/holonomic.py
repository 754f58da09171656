"""Holonomic vehicle: x(t), y(t) are degree-3 splines; velocity/acceleration
bounds are linear rows on the derivative coefficients (reference
``omgtools/vehicles/holonomic.py``: bounds 30-57, trajectory constraints 62-85,
initial/terminal constraints 87-105, initial guess 118-127, parameters 153-159,
collision constraints 161-163)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/holonomic.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.shape import Circle
from ..basics.spline_extra import sample_splines


class Holonomic(Vehicle):

    def __init__(self, shapes=None, options=None, bounds=None):
        bounds = bounds or {}
        shapes = shapes if shapes is not None else Circle(0.1)
        Vehicle.__init__(self, n_spl=2, degree=3, shapes=shapes, options=options)
        if self.options['syslimit'] == 'norm_inf':
            self.vxmin = bounds.get('vxmin', -0.5)
            self.vymin = bounds.get('vymin', -0.5)
            self.vxmax = bounds.get('vxmax', 0.5)
            self.vymax = bounds.get('vymax', 0.5)
            self.axmin = bounds.get('axmin', -1.)
            self.aymin = bounds.get('aymin', -1.)
            self.axmax = bounds.get('axmax', 1.)
            self.aymax = bounds.get('aymax', 1.)
            if 'vmin' in bounds:
                self.vxmin = self.vymin = bounds['vmin']
            if 'vmax' in bounds:
                self.vxmax = self.vymax = bounds['vmax']
            if 'amin' in bounds:
                self.axmin = self.aymin = bounds['amin']
            if 'amax' in bounds:
                self.axmax = self.aymax = bounds['amax']
        elif self.options['syslimit'] == 'norm_2':
            self.vmax = bounds.get('vmax', 0.5)
            self.amax = bounds.get('amax', 1.)

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options.update({'syslimit': 'norm_inf'})

    def define_trajectory_constraints(self, splines, horizon_time):
        x, y = splines
        dx, dy = x.derivative(), y.derivative()
        ddx, ddy = x.derivative(2), y.derivative(2)
        T = horizon_time
        if self.options['syslimit'] == 'norm_2':
            self.define_constraint(
                (dx**2 + dy**2) - (T**2) * self.vmax**2, -inf, 0.)
            self.define_constraint(
                (ddx**2 + ddy**2) - (T**4) * self.amax**2, -inf, 0.)
        elif self.options['syslimit'] == 'norm_inf':
            self.define_constraint(-dx + T * self.vxmin, -inf, 0.)
            self.define_constraint(-dy + T * self.vymin, -inf, 0.)
            self.define_constraint(dx - T * self.vxmax, -inf, 0.)
            self.define_constraint(dy - T * self.vymax, -inf, 0.)
            self.define_constraint(-ddx + (T**2) * self.axmin, -inf, 0.)
            self.define_constraint(-ddy + (T**2) * self.aymin, -inf, 0.)
            self.define_constraint(ddx - (T**2) * self.axmax, -inf, 0.)
            self.define_constraint(ddy - (T**2) * self.aymax, -inf, 0.)
        else:
            raise ValueError(
                'Only norm_2 and norm_inf are defined as system limit.')

    def get_initial_constraints(self, splines, horizon_time):
        state0 = self.define_parameter('state0', 2)
        input0 = self.define_parameter('input0', 2)
        x, y = splines
        dx, dy = x.derivative(), y.derivative()
        return [(x, state0[0]), (y, state0[1]),
                (dx, horizon_time * input0[0]), (dy, horizon_time * input0[1])]

    def get_terminal_constraints(self, splines, horizon_time=None):
        position = self.define_parameter('poseT', 2)
        x, y = splines
        term_con = [(x, position[0]), (y, position[1])]
        term_con_der = []
        for d in range(1, self.degree + 1):
            term_con_der.extend([(x.derivative(d), 0.), (y.derivative(d), 0.)])
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.zeros(2)
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)
        self.prediction['dinput'] = np.zeros(2)

    def set_terminal_conditions(self, position):
        self.poseT = np.asarray(position, dtype=float)

    def get_init_spline_value(self, subgoals=None):
        pos0, posT = self.prediction['state'], self.poseT
        init_value = np.zeros((len(self.basis), 2))
        for k in range(2):
            init_value[:, k] = np.linspace(pos0[k], posT[k], len(self.basis))
        return [init_value]

    def check_terminal_conditions(self):
        tol = self.options['stop_tol']
        if (np.linalg.norm(self.signals['state'][:, -1] - self.poseT) > tol or
                np.linalg.norm(self.signals['input'][:, -1]) > tol):
            return False
        return True

    def set_parameters(self, current_time):
        parameters = Vehicle.set_parameters(self, current_time)
        parameters[self]['state0'] = self.prediction['state']
        parameters[self]['input0'] = self.prediction['input']
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
        inp = np.c_[sample_splines([dx, dy], time)]
        signals['state'] = np.c_[sample_splines([x, y], time)]
        signals['input'] = inp
        signals['v_tot'] = np.sqrt(inp[0, :]**2 + inp[1, :]**2)
        signals['dinput'] = np.c_[sample_splines([ddx, ddy], time)]
        return signals

    def state2pose(self, state):
        return np.r_[state, 0.]

    def ode(self, state, input):
        return input
