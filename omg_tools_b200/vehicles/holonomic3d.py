"""Holonomic vehicle in 3D: x(t), y(t), z(t) are degree-3 splines with box
(norm_inf) or Euclidean (norm_2) bounds on velocity and acceleration (reference
``omgtools/vehicles/holonomic3d.py``: bounds 29-37, trajectory constraints
46-74, initial/terminal constraints 76-96, initial guess 107-115, parameters
126-131, 3D collision rows 133-135)."""
# Attribution: the class / method / option names and the constraint rows of this module restate
# the corresponding module of OMG-tools (omgtools/vehicles/holonomic3d.py; Copyright (C) 2016 Ruben Van Parys &
# Tim Mercy, KU Leuven; GNU LGPL v3) -- they are the drop-in contract of this framework.  See NOTICE.
import numpy as np

from .vehicle import Vehicle
from ..basics.optilayer import inf
from ..basics.spline_extra import sample_splines


class Holonomic3D(Vehicle):

    def __init__(self, shapes, options=None, bounds=None):
        bounds = bounds or {}
        Vehicle.__init__(self, n_spl=3, degree=3, shapes=shapes, options=options)
        self.vmin = bounds.get('vmin', -0.5)
        self.vmax = bounds.get('vmax', 0.5)
        self.amin = bounds.get('amin', -1.)
        self.amax = bounds.get('amax', 1.)

    def set_default_options(self):
        Vehicle.set_default_options(self)
        self.options.update({'syslimit': 'norm_inf'})

    def define_trajectory_constraints(self, splines, horizon_time):
        T = horizon_time
        vel = [s.derivative() for s in splines]
        acc = [s.derivative(2) for s in splines]
        if self.options['syslimit'] == 'norm_2':
            self.define_constraint(
                sum(d**2 for d in vel) - (T**2) * self.vmax**2, -inf, 0.)
            self.define_constraint(
                sum(d**2 for d in acc) - (T**4) * self.amax**2, -inf, 0.)
        elif self.options['syslimit'] == 'norm_inf':
            # same row order as the reference: all lower velocity rows, all
            # upper velocity rows, then the same for the acceleration
            for d in vel:
                self.define_constraint(-d + T * self.vmin, -inf, 0.)
            for d in vel:
                self.define_constraint(d - T * self.vmax, -inf, 0.)
            for d in acc:
                self.define_constraint(-d + (T**2) * self.amin, -inf, 0.)
            for d in acc:
                self.define_constraint(d - (T**2) * self.amax, -inf, 0.)
        else:
            raise ValueError(
                'Only norm_2 and norm_inf are defined as system limit.')

    def get_initial_constraints(self, splines, horizon_time):
        state0 = self.define_parameter('state0', 3)
        input0 = self.define_parameter('input0', 3)
        con = [(s, state0[k]) for k, s in enumerate(splines)]
        con += [(s.derivative(), horizon_time * input0[k])
                for k, s in enumerate(splines)]
        return con

    def get_terminal_constraints(self, splines, horizon_time=None):
        position = self.define_parameter('poseT', 3)
        term_con = [(s, position[k]) for k, s in enumerate(splines)]
        term_con_der = []
        for d in range(1, self.degree + 1):
            term_con_der.extend([(s.derivative(d), 0.) for s in splines])
        return [term_con, term_con_der]

    def set_initial_conditions(self, state, input=None):
        if input is None:
            input = np.zeros(3)
        self.prediction['state'] = np.asarray(state, dtype=float)
        self.prediction['input'] = np.asarray(input, dtype=float)

    def set_terminal_conditions(self, position):
        self.poseT = np.asarray(position, dtype=float)

    def get_init_spline_value(self, subgoals=None):
        pos0, posT = self.prediction['state'], self.poseT
        init_value = np.zeros((len(self.basis), 3))
        for k in range(3):
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
        self.define_collision_constraints_3d(
            hyperplanes, room, [splines[0], splines[1], splines[2]], horizon_time)

    def splines2signals(self, splines, time):
        signals = {}
        vel = [s.derivative() for s in splines[:3]]
        acc = [s.derivative(2) for s in splines[:3]]
        inp = np.c_[sample_splines(vel, time)]
        signals['state'] = np.c_[sample_splines(list(splines[:3]), time)]
        signals['input'] = inp
        signals['v_tot'] = np.sqrt(inp[0, :]**2 + inp[1, :]**2 + inp[2, :]**2)
        signals['a'] = np.c_[sample_splines(acc, time)]
        return signals

    def state2pose(self, state):
        return np.r_[state, np.zeros(3)]

    def ode(self, state, input):
        return input
